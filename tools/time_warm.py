"""Warm batches (records one tick later, carrying the previous tick's working sets): one-launch kernel against the three
passes with early finish, device-resident, microseconds per call for several batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import OUT_DTYPE, default_params, lib, states
sizes = [16384, 65536, 262144, 1048576]
nmax = max(sizes)
S = states.generate_states(nmax, 20260103, masks="mixed")
d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(nmax * 256, dtype=torch.uint8, device="cuda")
sol = lib.BalanceSolver(default_params(0.6))
sol.control_packed(d_in, d_out, nmax)
torch.cuda.synchronize()
rng = np.random.default_rng(3)
S["pad"][:, :4] = d_out.cpu().numpy().view(OUT_DTYPE)["pad"][:, :4]
S["x"] += rng.normal(0, 2e-4, S["x"].shape)
S["xdot"] += rng.normal(0, 2e-3, S["xdot"].shape)
S["w"] += rng.normal(0, 2e-3, S["w"].shape)
d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
sol.close()
stream = torch.cuda.current_stream()
print("mode            " + "".join(f"{n:>10d}" for n in sizes) + "   (us per call; mixed-contact states one tick later)")
for name, env in (("one launch", str(1 << 40)), ("early finish", "0")):
    os.environ["QPB_TPQ_WARM_DEFER_MIN"] = env
    sol = lib.BalanceSolver(default_params(0.6))
    sol.set_warm_batches(True)
    row = []
    for n in sizes:
        for _ in range(3):
            sol.control_packed(d_in, d_out, n, stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            sol.control_packed(d_in, d_out, n, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) * 1e2)
    it = d_out.cpu().numpy().view(OUT_DTYPE)["iters"]
    print(f"{name:16s}" + "".join(f"{t:10.1f}" for t in row) + f"   iters mean {it.mean():.3f} max {it.max()}")
    sol.close()
