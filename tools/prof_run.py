"""Tiny driver for ncu: a few launches of the packed balance kernel on one BASELINE workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import lib, states, default_params
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n, seed, masks, profile, desc = WORKLOADS[wl]
sol = lib.BalanceSolver(default_params(0.6))
bufs = [torch.from_numpy(states.generate_states(n, seed + 1000 * k, profile=profile, masks=masks).view(np.uint8).reshape(-1)).cuda() for k in range(2)]
out = torch.empty(n * 256, dtype=torch.uint8, device="cuda")
for i in range(reps):
    sol.control_packed(bufs[i % 2], out, n)
torch.cuda.synchronize()
print("done", wl, reps)
