"""Times the packed balance kernel of the library named by $QPB_LIB on one workload (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from quadruped_control_b200 import lib, states, default_params, OUT_DTYPE
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n, seed, masks, profile, desc = WORKLOADS[wl]
if len(sys.argv) > 2: n = int(sys.argv[2])
sol = lib.BalanceSolver(default_params(0.6))
nb = 8 if n <= 131072 else 2
hb = [states.generate_states(n, seed + 1000 * k, profile=profile, masks=masks) for k in range(nb)]
bufs = [torch.from_numpy(b.view(np.uint8).reshape(-1)).cuda() for b in hb]
outs = [torch.empty(n * 256, dtype=torch.uint8, device="cuda") for _ in range(nb)]
for i in range(5): sol.control_packed(bufs[i % nb], outs[i % nb], n)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): sol.control_packed(bufs[i % nb], outs[i % nb], n)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 40)
o = outs[7 % nb].cpu().numpy().view(OUT_DTYPE)
import oracle
idx = np.arange(0, n, max(1, n // 1024))
ref = oracle.control_batch(default_params(0.6), np.ascontiguousarray(hb[7 % nb][idx]), 8)
err = float((np.abs(o["grf_body"][idx] - ref["grf_body"]).max(axis=1) / np.maximum(np.abs(ref["grf_body"]).max(axis=1), 1)).max())
print(f"{os.environ.get('QPB_LIB','default'):40s} {wl} n={n} ms={best:.4f} QP/s={n/best*1e3:.4e} err={err:.2e} bad={int((o['status']!=0).sum())}")
