"""Developer tool (run under ncu): two launches each of the planner, adapter, swing and torque-command kernels on
1 048 576 mixed-contact robots, so one capture shows what bounds each of them."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_control_b200 import default_params, lib, states  # noqa: E402
from quadruped_control_b200.records import (COM_MSG_DTYPE, JOINT_MSG_DTYPE, OUT_DTYPE, PLAN_DTYPE, SWING_DTYPE,  # noqa: E402
                                            TORQUE_CMD_DTYPE)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
s = lib.BalanceSolver(default_params(0.6), device=0)
S = states.generate_states(n, 20260103, masks="mixed")
SW = states.generate_swing(S, 5)
plan = np.zeros(n, dtype=PLAN_DTYPE)
plan["phase"] = np.random.default_rng(0).uniform(0.8, 1.0, size=(n, 4))
plan["replan"] = 1


def dev(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()


d_S, d_plan, d_sw = dev(S), dev(plan), dev(SW)
d_out = torch.zeros(n * OUT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_cmd = torch.zeros(n * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_com = torch.zeros(n * COM_MSG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_com.view(torch.float64).view(n, 13)[:, 6] = 1.0
d_js = torch.zeros(n * JOINT_MSG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
s.plan(d_S, d_plan, d_sw, n, stream=st)       # re-plans every swing leg
s.plan(d_S, d_plan, d_sw, n, stream=st)       # references only
d_sw2 = dev(SW)
s.tick_packed(d_S, d_sw2, d_out, n, stream=st)  # balance + swing
s.tick_packed(d_S, d_sw2, d_out, n, stream=st)
s.torque_cmd(d_S, d_out, d_cmd, n, stream=st)
s.torque_cmd(d_S, d_out, d_cmd, n, stream=st)
s.adapt_inputs(d_com, d_js, d_S, d_sw, n, stream=st)
s.adapt_inputs(d_com, d_js, d_S, d_sw, n, stream=st)
torch.cuda.synchronize()
