"""Developer tool: time the planner / adapter / torque-command kernels on 1 048 576 mixed-contact robots
(CUDA events, device-resident records) and print achieved HBM traffic."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_control_b200 import default_params, lib, states  # noqa: E402
from quadruped_control_b200.records import (COM_MSG_DTYPE, JOINT_MSG_DTYPE, OUT_DTYPE, PLAN_DTYPE, STATE_DTYPE, SWING_DTYPE,  # noqa: E402
                                            TORQUE_CMD_DTYPE)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
params = default_params(0.6)
s = lib.BalanceSolver(params, device=0)
S = states.generate_states(n, 20260103, masks="mixed")
rng = np.random.default_rng(0)
plan = np.zeros(n, dtype=PLAN_DTYPE)
plan["phase"] = rng.uniform(0.8, 1.0, size=(n, 4))
plan["replan"] = 1


def dev(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()


d_S, d_plan0 = dev(S), dev(plan)
d_plan = d_plan0.clone()
d_sw = torch.zeros(n * SWING_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n * OUT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_cmd = torch.zeros(n * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_com = torch.zeros(n * COM_MSG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_com.view(torch.float64).view(n, 13)[:, 6] = 1.0  # unit quaternion
d_js = torch.zeros(n * JOINT_MSG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
swing_legs = int((S["contact"] == 0).sum())


def timed(name, fn, nbytes, reps=10, setup=None):
    for _ in range(2):
        if setup:
            setup()
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if setup:
            setup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    print(f"{name}: {ms:.3f} ms per {n} robots = {n / ms * 1e3:.3g} robots/s, {nbytes / ms / 1e6:.0f} GB/s algorithmic")


# bytes: what the kernel must read and write per robot
timed("plan_kernel (all swing legs re-planned)", lambda: s.plan(d_S, d_plan, d_sw, n, stream=st),
      n * (21 * 8 + 4 + 32 + 4) + swing_legs * (3 * 8 + 2 * 24 + 2 * 24) + 0 * n, setup=lambda: d_plan.copy_(d_plan0))
timed("plan_kernel (no re-plan)", lambda: s.plan(d_S, d_plan, d_sw, n, stream=st), n * (12 * 8 + 4 + 32 + 4) + swing_legs * (4 * 24))
timed("adapt_kernel", lambda: s.adapt_inputs(d_com, d_js, d_S, d_sw, n, stream=st), n * (104 + 192 + 42 * 8 + 96))
timed("torque_cmd_kernel", lambda: s.torque_cmd(d_S, d_out, d_cmd, n, stream=st), n * (4 + 96 + 4 + 112))
