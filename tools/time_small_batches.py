"""Latency of small batches: CUDA-event time per launch of qpb_control_batch_packed for n = 1 ... 65536 records, for the
one-launch half-warp kernel, the range-space path with 1 / 2 / 4 lanes per QP (QPB_TPQ_MIN_N=0 forces it at every size) and
the one-launch range-space kernel (QPB_TPQ_ONE_MAX), cold and with every record carrying its own final working set as the hint.
Tells where the dispatch threshold (QPB_TPQ_MIN_N) belongs and what a loop iteration costs when the machine is empty."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quadruped_control_b200 import OUT_DTYPE, default_params, lib, states

sizes = [1, 32, 148, 1024, 4096, 8192, 16384, 32768, 65536]
S = states.generate_states(max(sizes), 20260102)
d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(len(S) * 256, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
print("mode        " + "".join(f"{n:>10d}" for n in sizes) + "   (us per launch)")
# hinted copy of the records: every state carries the working set its own cold solve ended on (a perfect warm start)
os.environ["QPB_TPQ_MIN_N"] = "0"
sol = lib.BalanceSolver(default_params(0.6))
sol.control_packed(d_in, d_out, len(S), stream.cuda_stream)
torch.cuda.synchronize()
W = S.copy()
W["pad"][:, :4] = d_out.cpu().numpy().view(OUT_DTYPE)["pad"][:, :4]
d_warm = torch.from_numpy(W.view(np.uint8).reshape(-1)).cuda()
sol.close()
ONE = {"QPB_TPQ_ONE_MAX": str(1 << 30), "QPB_TPQ_MIN_N": str(1 << 31)}
for name, env, d_in in (("half-warp", {"QPB_QPS_PER_WARP": "2"}, d_in), ("range lpq1", {"QPB_TPQ_MIN_N": "0", "QPB_TPQ_LPQ": "1"}, d_in),
                        ("range lpq2", {"QPB_TPQ_MIN_N": "0", "QPB_TPQ_LPQ": "2"}, d_in),
                        ("range lpq4", {"QPB_TPQ_MIN_N": "0", "QPB_TPQ_LPQ": "4"}, d_in), ("one-launch", ONE, d_in),
                        ("lpq1 warm", {"QPB_TPQ_MIN_N": "0", "QPB_TPQ_LPQ": "1"}, d_warm), ("one warm", ONE, d_warm)):
    for k in ("QPB_QPS_PER_WARP", "QPB_TPQ_MIN_N", "QPB_TPQ_LPQ", "QPB_TPQ_ONE_MAX"):
        os.environ.pop(k, None)
    os.environ.update(env)
    sol = lib.BalanceSolver(default_params(0.6))
    row = []
    for n in sizes:
        for _ in range(3):
            sol.control_packed(d_in, d_out, n, stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for _ in range(reps):
            sol.control_packed(d_in, d_out, n, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) * 1e3 / reps)
    it = d_out.cpu().numpy().view(OUT_DTYPE)["iters"]
    print(f"{name:12s}" + "".join(f"{t:10.1f}" for t in row) + f"   iters mean {it.mean():.1f} max {it.max()}")
    sol.close()
