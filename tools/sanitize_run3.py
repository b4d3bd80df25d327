"""Small run of the 256-bit record kernels (both alignments, full tiles and the tail) and the single-process
multi-device calls for compute-sanitizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quadruped_control_b200 import default_params, lib, states
from quadruped_control_b200.records import COM_MSG_DTYPE, JOINT_MSG_DTYPE, OUT_DTYPE, PLAN_DTYPE, SWING_DTYPE, TORQUE_CMD_DTYPE


def dev(a, shift=0):
    raw = torch.zeros(a.nbytes + shift, dtype=torch.uint8, device="cuda")
    raw[shift:].copy_(torch.from_numpy(a.view(np.uint8).reshape(-1).copy()))
    return raw[shift:]


params = default_params(0.6)
sol = lib.BalanceSolver(params)
for m, shift in ((1500, 0), (1500, 16), (19, 0)):
    S = states.generate_states(m, 9, masks="mixed")
    plan = np.zeros(m, dtype=PLAN_DTYPE)
    plan["phase"] = np.random.default_rng(0).uniform(0.7, 1.0, size=(m, 4))
    plan["replan"] = 1
    com = np.zeros(m, dtype=COM_MSG_DTYPE)
    com["orientation"][:, 3] = 1.0
    js = np.zeros(m, dtype=JOINT_MSG_DTYPE)
    js["position"] = np.tile(states.STANCE_Q.reshape(4, 3).T.reshape(12), (m, 1))
    d_S, d_plan = dev(S, shift), dev(plan, shift)
    d_sw = dev(np.zeros(m, dtype=SWING_DTYPE), shift)
    d_out = dev(np.zeros(m, dtype=OUT_DTYPE), shift)
    d_cmd = dev(np.zeros(m, dtype=TORQUE_CMD_DTYPE), shift)
    sol.adapt_inputs(dev(com, shift), dev(js, shift), d_S, d_sw, m)
    sol.plan(d_S, d_plan, d_sw, m)
    sol.tick_packed(d_S, d_sw, d_out, m)
    sol.torque_cmd(d_S, d_out, d_cmd, m)
    torch.cuda.synchronize()
multi = lib.MultiBalanceSolver(params, devices=list(range(torch.cuda.device_count())) + [0])
S = states.generate_states(701, 3, masks="mixed")
a = multi.control_host(S)
b = sol.control_host(S)
assert a.tobytes() == b.tobytes()
multi.close()
sol.close()
print("sanitize_run3 ok")
