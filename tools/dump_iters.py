import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quadruped_control_b200 import lib, states, default_params
os.environ["QPB_QPS_PER_WARP"] = "1"
S = states.generate_states(200000, 20260103, masks="mixed")
sol = lib.BalanceSolver(default_params(0.6)); out = sol.control_host(S)
np.savez_compressed("gpurun_out/iters_cfg3.npz", iters=out["iters"], nst=S["contact"].sum(axis=1).astype(np.int8))
print("dumped", out["iters"].mean())
