"""One small run of EVERY kernel family for compute-sanitizer (tests/test_sanitizer_gpu.py drives it with memcheck and
racecheck): the range-space path (set-up / loop / finish, one, two and four lanes per QP, cold and warm-started, ragged
sizes), the half-warp and one-warp-per-QP kernels, the argument-per-array entry point, the asynchronous host pipeline,
the whole tick (swing legs), planner / adapters / torque command with both record alignments, the MPC kernel, and the
single-process multi-device calls.  Batches are a few hundred records: the sanitizer runs kernels ~50x slower."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quadruped_control_b200 import OUT_DTYPE, STATE_DTYPE, default_params, lib, states
from quadruped_control_b200.records import (COM_MSG_DTYPE, JOINT_MSG_DTYPE, PLAN_DTYPE, SWING_DTYPE, TORQUE_CMD_DTYPE,
                                            default_mpc_params)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 701
os.environ["QPB_TPQ_MIN_N"] = "0"  # every batch size through the range-space kernels
params = default_params(0.6)
S = states.generate_states(n, 4, profile="stress", masks="mixed")
S["contact"][:64] = (np.arange(64)[:, None] % 16 >> np.arange(4)) & 1
S["x"][3, 0] = np.nan
S["q"][5, 2] = np.inf


def dev(a, shift=0):
    raw = torch.zeros(a.nbytes + shift, dtype=torch.uint8, device="cuda")
    raw[shift:].copy_(torch.from_numpy(a.view(np.uint8).reshape(-1).copy()))
    return raw[shift:]


ref = None
# (kernel mapping, lanes per QP of the loop kernel, programmatic dependent launches of the second and third pass)
for mode, lpq, pdl in (("32", "1", "1"), ("32", "1", "0"), ("32", "2", "1"), ("32", "4", "1"), ("2", "1", "1"), ("1", "1", "1")):
    os.environ["QPB_QPS_PER_WARP"], os.environ["QPB_TPQ_LPQ"], os.environ["QPB_TPQ_PDL"] = mode, lpq, pdl
    sol = lib.BalanceSolver(params)
    out = sol.control_host(S)
    assert list(np.nonzero(out["status"])[0]) == [3, 5]
    if ref is None:
        ref = out
    assert np.abs(out["grf_body"] - ref["grf_body"]).max() < 1e-6
    d_in, d_out = dev(S), torch.empty(n * 256, dtype=torch.uint8, device="cuda")
    sol.control_packed(d_in, d_out, n)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().view(OUT_DTYPE).tobytes() == out.tobytes()
    W = S.copy()  # warm start: last tick's working sets (zero words from the one-launch kernels mean "no hint")
    W["pad"][:, :4] = out["pad"][:, :4]
    warm = sol.control_host(W)
    assert np.abs(warm["grf_body"] - ref["grf_body"]).max() < 1e-6
    d = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in ("Rwb", "Rwb_d", "x", "xdot", "w", "x_d", "xdot_d", "w_d", "feet", "contact", "q")}
    grf = torch.zeros(n, 12, dtype=torch.float64, device="cuda")
    tau, st = torch.zeros_like(grf), torch.zeros(n, dtype=torch.int32, device="cuda")
    sol.control_split(n, d["Rwb"], d["Rwb_d"], d["x"], d["xdot"], d["w"], d["x_d"], d["xdot_d"], d["w_d"], d["feet"], d["contact"], d["q"], grf, tau, st)
    pin_i, pin_o = lib.PinnedBuffer(n, STATE_DTYPE), [lib.PinnedBuffer(n, OUT_DTYPE) for _ in range(2)]
    pin_i.array[:] = S
    for k in range(4):
        sol.control_host_async(pin_i.array, pin_o[k % 2].array)
    sol.host_sync()
    assert pin_o[0].array.tobytes() == out.tobytes() and pin_o[1].array.tobytes() == out.tobytes()
    sol.fk_host(S["q"][:64])
    sol.jt_host(S["q"][:64], out["grf_body"][:64], S["contact"][:64])
    sw = states.generate_swing(S, 5, params)
    sol.tick_host(S, sw)
    for b in [pin_i] + pin_o:
        b.free()
    sol.close()
del os.environ["QPB_QPS_PER_WARP"], os.environ["QPB_TPQ_LPQ"], os.environ["QPB_TPQ_PDL"]

sol = lib.BalanceSolver(params)
for m, shift in ((300, 0), (300, 16), (19, 0)):  # record kernels: both alignments, full tiles and the tail
    Sm = states.generate_states(m, 9, masks="mixed")
    plan = np.zeros(m, dtype=PLAN_DTYPE)
    plan["phase"] = np.random.default_rng(0).uniform(0.7, 1.0, size=(m, 4))
    plan["replan"] = 1
    com = np.zeros(m, dtype=COM_MSG_DTYPE)
    com["orientation"][:, 3] = 1.0
    js = np.zeros(m, dtype=JOINT_MSG_DTYPE)
    js["position"] = np.tile(states.STANCE_Q.reshape(4, 3).T.reshape(12), (m, 1))
    d_S, d_plan = dev(Sm, shift), dev(plan, shift)
    d_sw = dev(np.zeros(m, dtype=SWING_DTYPE), shift)
    d_o = dev(np.zeros(m, dtype=OUT_DTYPE), shift)
    d_cmd = dev(np.zeros(m, dtype=TORQUE_CMD_DTYPE), shift)
    sol.adapt_inputs(dev(com, shift), dev(js, shift), d_S, d_sw, m)
    sol.plan(d_S, d_plan, d_sw, m)
    sol.tick_packed(d_S, d_sw, d_o, m)
    sol.torque_cmd(d_S, d_o, d_cmd, m)
    torch.cuda.synchronize()
multi = lib.MultiBalanceSolver(params, devices=list(range(torch.cuda.device_count())) + [0])
a, b = multi.control_host(S), sol.control_host(S)
assert a.tobytes() == b.tobytes()
multi.close()
sol.close()

R = np.concatenate([states.generate_mpc(48, 3), states.generate_mpc(16, 4, scale=4.0)])
R["contact"][0] = 0
R["x0"][1, 3] = np.nan
mpc = lib.MpcSolver(default_mpc_params(0.6))
mo = mpc.solve_host(R)
assert set(np.unique(mo["status"])) <= {0, 2}
mpc.close()
print("sanitize_run ok", n)
