"""Tiny driver for ncu: the one-launch range-space kernel on n warm records (each carries its own final working set).
usage: python tools/prof_one.py [n] [reps] [cold]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import OUT_DTYPE, default_params, lib, states
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cold = len(sys.argv) > 3
S = states.generate_states(n, 20260102)
sol = lib.BalanceSolver(default_params(0.6))
out = torch.empty(n * 256, dtype=torch.uint8, device="cuda")
if not cold:
    sol.control_packed(torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda(), out, n)
    torch.cuda.synchronize()
    S["pad"][:, :4] = out.cpu().numpy().view(OUT_DTYPE)["pad"][:, :4]
d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
sol.set_warm_batches(True)
for _ in range(reps):
    sol.control_packed(d_in, out, n)
torch.cuda.synchronize()
print("done", n, reps, "iters", out.cpu().numpy().view(OUT_DTYPE)["iters"].mean())
