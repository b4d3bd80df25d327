"""Small run of the MPC, planner, adapter and torque-command entry points for compute-sanitizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quadruped_control_b200 import default_params, lib, states
from quadruped_control_b200.records import (COM_MSG_DTYPE, JOINT_MSG_DTYPE, OUT_DTYPE, PLAN_DTYPE, SWING_DTYPE, TORQUE_CMD_DTYPE,
                                            default_mpc_params)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
R = np.concatenate([states.generate_mpc(n, 3), states.generate_mpc(n // 2, 4, scale=4.0)])
R["contact"][0] = 0
R["x0"][1, 3] = np.nan
mpc = lib.MpcSolver(default_mpc_params(0.6))
out = mpc.solve_host(R)
assert set(np.unique(out["status"])) <= {0, 2}
mpc.close()


def dev(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()


m = 1500
sol = lib.BalanceSolver(default_params(0.6))
S = states.generate_states(m, 9, masks="mixed")
plan = np.zeros(m, dtype=PLAN_DTYPE)
plan["phase"] = np.random.default_rng(0).uniform(0.7, 1.0, size=(m, 4))
plan["replan"] = 1
com = np.zeros(m, dtype=COM_MSG_DTYPE)
com["orientation"][:, 3] = 1.0
js = np.zeros(m, dtype=JOINT_MSG_DTYPE)
js["position"] = np.tile(states.STANCE_Q.reshape(4, 3).T.reshape(12), (m, 1))
d_S, d_plan = dev(S), dev(plan)
d_sw = torch.zeros(m * SWING_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(m * OUT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_cmd = torch.zeros(m * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
sol.adapt_inputs(dev(com), dev(js), d_S, d_sw, m)
sol.plan(d_S, d_plan, d_sw, m)
sol.tick_packed(d_S, d_sw, d_out, m)
sol.torque_cmd(d_S, d_out, d_cmd, m)
torch.cuda.synchronize()
sol.close()
print("sanitize_run2 ok", len(R), m)
