"""Foothold planner + swing-foot trajectory and message adapters (rank 3).

The oracle port (oracle/plan_oracle.c) is pinned against the reference's own foot_planner.cpp and trajectory.cpp compiled
in oracle/_ref (stand-in Armadillo / ROS console / rigid3d headers); the CUDA kernels are compared with the oracle on
the GPU.  Tolerance: 1e-9 relative (floor 1) for the trajectory -- the kernel evaluates the closed-form solution of the
7x7 system the reference solves by LU every time -- and exact for pure data movement."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_err
from quadruped_control_b200 import default_params, lib, states
from quadruped_control_b200.records import (COM_MSG_DTYPE, JOINT_MSG_DTYPE, OUT_DTYPE, PLAN_DTYPE, STATE_DTYPE, SWING_DTYPE,
                                            TORQUE_CMD_DTYPE, PlanParams, default_joint_gains, default_plan_params)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_plan(S, seed, all_replan=False):
    rng = np.random.default_rng(seed)
    n = len(S)
    plan = np.zeros(n, dtype=PLAN_DTYPE)
    plan["phase"] = rng.uniform(0.6, 1.05, size=(n, 4))  # stance_phase = 0.816: some phases clamp to t = 0, some to t = 1
    plan["replan"] = 1 if all_replan else rng.integers(0, 2, size=(n, 4))
    plan["p_start"] = rng.normal(0, 0.3, size=(n, 12))
    plan["p_final"] = rng.normal(0, 0.3, size=(n, 12))
    return plan


def make_msgs(n, seed):
    rng = np.random.default_rng(seed)
    com = np.zeros(n, dtype=COM_MSG_DTYPE)
    com["position"] = rng.normal(0, 0.5, size=(n, 3)) + [0, 0, 0.26]
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[::7] *= 1.003  # slightly non-unit quaternions: the reference's conversion divides by |q|^2 instead of normalising
    com["orientation"] = q
    com["linear"] = rng.normal(0, 0.3, size=(n, 3))
    com["angular"] = rng.normal(0, 0.5, size=(n, 3))
    js = np.zeros(n, dtype=JOINT_MSG_DTYPE)
    js["position"] = rng.uniform(-1.5, 1.5, size=(n, 12))
    js["velocity"] = rng.normal(0, 2.0, size=(n, 12))
    return com, js


# ---------------------------------------------------------------------------------------------------- CPU --
def test_plan_layouts_and_defaults(built, tmp_path):
    import oracle

    assert bytes(lib_default_plan()) == bytes(default_plan_params()) == bytes(oracle.plan_default_params())
    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "qpb200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(qpb_plan_params), sizeof(qpb_plan_rec), sizeof(qpb_com_msg),'
        " sizeof(qpb_joint_msg), sizeof(qpb_torque_cmd), offsetof(qpb_plan_rec, phase), offsetof(qpb_plan_rec, replan),"
        " offsetof(qpb_torque_cmd, count));return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [ctypes.sizeof(PlanParams), PLAN_DTYPE.itemsize, COM_MSG_DTYPE.itemsize, JOINT_MSG_DTYPE.itemsize,
                   TORQUE_CMD_DTYPE.itemsize, PLAN_DTYPE.fields["phase"][1], PLAN_DTYPE.fields["replan"][1],
                   TORQUE_CMD_DTYPE.fields["count"][1]]


def lib_default_plan():
    p = PlanParams()
    assert lib.load().qpb_default_plan_params(ctypes.byref(p)) == 0
    return p


def test_plan_oracle_matches_reference_sources(built, params06):
    """oracle/plan_oracle.c against FootPlanner::positions/singleFoot and FootTrajectoryManager::referenceStates of the
    reference's own foot_planner.cpp / trajectory.cpp (oracle/_ref)."""
    import oracle

    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    pp = default_plan_params()
    for seed, masks in ((1, "mixed"), (2, "all4")):
        S = states.generate_states(1500, seed, masks=masks)
        if masks == "all4":
            S["contact"][::3] = 0  # flight: all four legs in swing
        plan_a, plan_b = make_plan(S, seed), make_plan(S, seed)
        sw_a, sw_b = np.zeros(len(S), dtype=SWING_DTYPE), np.zeros(len(S), dtype=SWING_DTYPE)
        oracle.plan_batch(pp, S, plan_a, sw_a)
        oracle.ref_plan_batch(pp, S, plan_b, sw_b)
        assert plan_a.tobytes() == plan_b.tobytes()
        assert np.abs(sw_a["foot_ref_pos"] - sw_b["foot_ref_pos"]).max() <= 1e-13
        assert np.abs(sw_a["foot_ref_vel"] - sw_b["foot_ref_vel"]).max() <= 1e-12
        swing = np.repeat(S["contact"] == 0, 3, axis=1)
        assert not sw_a["foot_ref_pos"][~swing].any() and sw_a["foot_ref_pos"][swing].any()
    for leg in range(4):
        assert np.array_equal(oracle.single_foot(pp, leg, S[5]), oracle.ref_single_foot(pp.t_stance, leg, S[5]))


def test_replanning_another_leg_keeps_a_swinging_leg_on_its_trajectory(built):
    """A deliberate deviation from the reference, pinned here.  FootTrajectoryManager::referenceStates(gait_map, bounds)
    clears every stored trajectory before adding the newly planned legs (trajectory.cpp:316), so when leg b re-plans while
    leg a is mid-swing the reference loses leg a's trajectory: referenceState(a) logs "Failed to find trajectory" and
    returns a zero FootState (trajectory.cpp:366-388), i.e. the swing foot is sent towards the world origin.  With the
    gaits the reference ships this never happens (trot: both legs of a pair re-plan in the same tick; crawl: one leg
    swings at a time).  The batched planner keeps p_start / p_final of every swinging leg in the caller's qpb_plan_rec, so
    leg a continues on its sextic.  This test runs both through ONE manager of the reference (oracle/_ref) and through
    the oracle port, and documents the difference."""
    import oracle

    assert oracle.ref_available()
    pp = default_plan_params()
    S = states.generate_states(1, 77, masks="all4")
    S["contact"][0] = (1, 0, 0, 1)  # legs FL (1) and RR (2) in swing
    a_start, a_final = np.array([0.20, 0.13, 0.0]), np.array([0.31, 0.14, 0.0])
    b_start, b_final = np.array([-0.19, -0.12, 0.0]), np.array([-0.08, -0.13, 0.0])
    phase_a, phase_b = 0.93, 0.83
    ref_a, ref_b = oracle.ref_two_tick_replan(pp, 1, a_start, a_final, phase_a, 2, b_start, b_final, phase_b)
    assert not ref_a.any()  # the reference dropped leg a's trajectory
    plan = np.zeros(1, dtype=PLAN_DTYPE)
    plan["phase"][0] = (0.0, phase_a, phase_b, 0.0)
    plan["p_start"][0, 3:6], plan["p_final"][0, 3:6] = a_start, a_final  # planned on an earlier tick
    plan["p_start"][0, 6:9], plan["p_final"][0, 6:9] = b_start, b_final
    sw = np.zeros(1, dtype=SWING_DTYPE)
    oracle.plan_batch(pp, S, plan, sw)
    # leg b: identical to the reference; leg a: what the reference itself returns when its trajectory is not cleared
    assert np.abs(sw["foot_ref_pos"][0, 6:9] - ref_b[:3]).max() <= 1e-13 and np.abs(sw["foot_ref_vel"][0, 6:9] - ref_b[3:]).max() <= 1e-12
    keep_a, _ = oracle.ref_two_tick_replan(pp, 2, b_start, b_final, phase_b, 1, a_start, a_final, phase_a)  # a planned last: kept
    _, kept = oracle.ref_two_tick_replan(pp, 2, b_start, b_final, phase_b, 1, a_start, a_final, phase_a)
    assert np.abs(sw["foot_ref_pos"][0, 3:6] - kept[:3]).max() <= 1e-13 and np.abs(sw["foot_ref_vel"][0, 3:6] - kept[3:]).max() <= 1e-12
    assert sw["foot_ref_pos"][0, 3:6].any()


def test_plan_oracle_matches_golden_reference_outputs(built):
    """tests/golden/plan_golden.npz holds outputs of the reference's own planner sources (generated where /root/reference
    exists); the oracle port must reproduce them anywhere."""
    import oracle

    g = np.load(os.path.join(ROOT, "tests", "golden", "plan_golden.npz"))
    S = np.ascontiguousarray(g["states"]).view(STATE_DTYPE).reshape(-1)
    plan = np.ascontiguousarray(g["plan_in"]).view(PLAN_DTYPE).reshape(-1).copy()
    want = np.ascontiguousarray(g["plan_out"]).view(PLAN_DTYPE).reshape(-1)
    sw = np.zeros(len(S), dtype=SWING_DTYPE)
    oracle.plan_batch(default_plan_params(), S, plan, sw)
    assert plan.tobytes() == want.tobytes()
    assert np.abs(sw["foot_ref_pos"] - g["foot_ref_pos"]).max() <= 1e-13 and np.abs(sw["foot_ref_vel"] - g["foot_ref_vel"]).max() <= 1e-12


def test_trajectory_meets_its_own_constraints(built):
    """trajectory.cpp:256-296: s(0) = p0, s(1) = pf, s(1/2) = pc, zero velocity and acceleration at both ends; and the
    closed form the CUDA kernel uses equals the LU solution."""
    import oracle

    rng = np.random.default_rng(3)
    for _ in range(20):
        p0, pc, pf = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
        rc, coef = oracle.foot_trajectory(p0, pc, pf)
        assert rc == 0
        for t, want in ((0.0, p0), (1.0, pf), (0.5, pc)):
            pos, vel = oracle.track_trajectory(coef, t)
            assert np.abs(pos - want).max() <= 1e-12
            if t != 0.5:
                assert np.abs(vel).max() <= 1e-11
        D, C = pf - p0, pc - p0
        closed = np.array([p0, 0 * p0, 0 * p0, 64 * C - 22 * D, -192 * C + 81 * D, 192 * C - 90 * D, -64 * C + 32 * D])
        assert np.abs(closed - coef).max() <= 1e-11 * (1 + np.abs(coef).max())
        acc = lambda t: sum(k * (k - 1) * coef[k] * t ** (k - 2) for k in range(2, 7))  # noqa: E731
        assert np.abs(acc(0.0)).max() <= 1e-11 and np.abs(acc(1.0)).max() <= 1e-9


def test_adapter_oracle_against_independent_restatement(built, params06):
    """stateCallback / jointCallback (commander_node.cpp:127-187) restated a second way: scipy's quaternion -> matrix
    and the index table of the callback written out literally."""
    import oracle
    from scipy.spatial.transform import Rotation

    com, js = make_msgs(200, 4)
    S = np.zeros(200, dtype=STATE_DTYPE)
    S["x_d"] = 7.0  # fields the adapter must not touch
    sw = np.zeros(200, dtype=SWING_DTYPE)
    oracle.adapt_inputs(params06, com, js, S, sw)
    R = Rotation.from_quat(com["orientation"]).as_matrix()  # scipy takes x y z w and normalises
    assert np.abs(S["Rwb"].reshape(-1, 3, 3) - R).max() <= 1e-14
    assert np.array_equal(S["x"], com["position"]) and np.array_equal(S["xdot"], com["linear"]) and np.array_equal(S["w"], com["angular"])
    table = {"RL": (0, 4, 8), "FL": (1, 5, 9), "RR": (2, 6, 10), "FR": (3, 7, 11)}  # commander_node.cpp:131-164
    for leg, name in enumerate(("RL", "FL", "RR", "FR")):
        for j, idx in enumerate(table[name]):
            assert np.array_equal(S["q"][:, 3 * leg + j], js["position"][:, idx])
            assert np.array_equal(sw["qdot"][:, 3 * leg + j], js["velocity"][:, idx])
    assert np.abs(S["feet"] - states.forward_kinematics(S["q"], params06)).max() <= 1e-15
    assert (S["x_d"] == 7.0).all()


def test_torque_cmd_oracle_order_and_clamp(built, params06):
    import oracle

    tau = np.arange(12, dtype=np.float64) * 4.0 - 22.0  # -22 .. 22: both clamps bite
    n, torque, legs = oracle.torque_cmd(params06, tau, [1, 1, 1, 1])
    assert n == 12 and list(legs) == [1, 1, 1, 3, 3, 3, 0, 0, 0, 2, 2, 2]  # std::map order FL FR RL RR
    want = np.clip(np.concatenate([tau[3:6], tau[9:12], tau[0:3], tau[6:9]]), -20.0, 20.0)
    assert np.array_equal(torque, want)
    n, torque, legs = oracle.torque_cmd(params06, tau, [1, 0, 0, 1])
    assert n == 6 and list(legs[:6]) == [3, 3, 3, 0, 0, 0]


# ---------------------------------------------------------------------------------------------------- GPU --
def _dev(a):
    import torch

    return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()


def _host(t, dtype):
    return t.cpu().numpy().view(dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("masks,seed", [("mixed", 21), ("all4", 22)])
def test_plan_gpu_matches_oracle(solver06, masks, seed):
    import oracle
    import torch

    pp = default_plan_params()
    S = states.generate_states(20000, seed, masks=masks)
    if masks == "all4":
        S["contact"][::2] = 0
        S["contact"][1::4, 2] = 0
    plan = make_plan(S, seed)
    sw = np.zeros(len(S), dtype=SWING_DTYPE)
    sw["qdot"] = 3.0  # must survive
    d_S, d_plan, d_sw = _dev(S), _dev(plan), _dev(sw)
    solver06.plan(d_S, d_plan, d_sw, len(S), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got_plan, got_sw = _host(d_plan, PLAN_DTYPE), _host(d_sw, SWING_DTYPE)
    ref_plan, ref_sw = plan.copy(), sw.copy()
    oracle.plan_batch(pp, S, ref_plan, ref_sw)
    assert np.array_equal(got_plan["replan"], ref_plan["replan"]) and np.array_equal(got_plan["phase"], ref_plan["phase"])
    assert rel_err(got_plan["p_start"], ref_plan["p_start"]) <= 1e-12 and rel_err(got_plan["p_final"], ref_plan["p_final"]) <= 1e-12
    assert rel_err(got_sw["foot_ref_pos"], ref_sw["foot_ref_pos"]) <= 1e-9
    assert rel_err(got_sw["foot_ref_vel"], ref_sw["foot_ref_vel"]) <= 1e-9
    assert (got_sw["qdot"] == 3.0).all()


@pytest.mark.gpu
def test_plan_gpu_matches_golden_reference_outputs(solver06):
    import torch

    g = np.load(os.path.join(ROOT, "tests", "golden", "plan_golden.npz"))
    S = np.ascontiguousarray(g["states"]).view(STATE_DTYPE).reshape(-1)
    plan = np.ascontiguousarray(g["plan_in"]).view(PLAN_DTYPE).reshape(-1)
    want = np.ascontiguousarray(g["plan_out"]).view(PLAN_DTYPE).reshape(-1)
    d_S, d_plan = _dev(S), _dev(plan)
    d_sw = torch.zeros(len(S) * SWING_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    solver06.plan(d_S, d_plan, d_sw, len(S), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got, sw = _host(d_plan, PLAN_DTYPE), _host(d_sw, SWING_DTYPE)
    assert np.array_equal(got["replan"], want["replan"])
    assert rel_err(got["p_start"], want["p_start"]) <= 1e-12 and rel_err(got["p_final"], want["p_final"]) <= 1e-12
    assert rel_err(sw["foot_ref_pos"], g["foot_ref_pos"]) <= 1e-9 and rel_err(sw["foot_ref_vel"], g["foot_ref_vel"]) <= 1e-9


@pytest.mark.gpu
def test_adapt_and_torque_cmd_gpu_match_oracle(solver06, params06):
    import oracle
    import torch

    n = 5000
    com, js = make_msgs(n, 31)
    S0 = states.generate_states(n, 31, masks="mixed")
    sw0 = states.generate_swing(S0, 32, params06)
    d_S, d_sw = _dev(S0), _dev(sw0)
    st = torch.cuda.current_stream().cuda_stream
    solver06.adapt_inputs(_dev(com), _dev(js), d_S, d_sw, n, stream=st)
    torch.cuda.synchronize()
    S_ref, sw_ref = S0.copy(), sw0.copy()
    oracle.adapt_inputs(params06, com, js, S_ref, sw_ref)
    got_S, got_sw = _host(d_S, STATE_DTYPE), _host(d_sw, SWING_DTYPE)
    for f in ("x", "xdot", "w", "q", "x_d", "xdot_d", "w_d", "Rwb_d", "contact"):
        assert np.array_equal(got_S[f], S_ref[f]), f
    assert np.abs(got_S["Rwb"] - S_ref["Rwb"]).max() <= 1e-15 and np.abs(got_S["feet"] - S_ref["feet"]).max() <= 1e-15
    assert np.array_equal(got_sw["qdot"], sw_ref["qdot"]) and np.array_equal(got_sw["foot_ref_pos"], sw_ref["foot_ref_pos"])

    out = np.zeros(n, dtype=OUT_DTYPE)
    rng = np.random.default_rng(5)
    out["tau"] = rng.normal(0, 15.0, size=(n, 12))
    out["status"] = rng.integers(0, 2, size=n)
    d_cmd = torch.zeros(n * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    solver06.torque_cmd(d_S, _dev(out), d_cmd, n, stream=st)
    torch.cuda.synchronize()
    cmd = _host(d_cmd, TORQUE_CMD_DTYPE)
    for i in range(0, n, 37):
        present = [(0 if (S0["contact"][i, leg] and out["status"][i] != 0) else 1) for leg in range(4)]
        cnt, torque, legs = oracle.torque_cmd(params06, out["tau"][i], present)
        assert cmd["count"][i] == cnt and np.array_equal(cmd["torque"][i][:cnt], torque[:cnt])
        assert list(cmd["leg"][i][:cnt]) == list(legs[:cnt]) and not cmd["torque"][i][cnt:].any()

    # Arrays that are only 16-byte aligned (the ABI minimum) take the double2 kernels instead of the 256-bit ones:
    # same bytes out.  A short batch (n < 32) has no full tile: the per-thread tail path of the 256-bit kernels.
    def shifted(a):
        raw = torch.zeros(a.nbytes + 16, dtype=torch.uint8, device="cuda")
        raw[16:].copy_(torch.from_numpy(a.view(np.uint8).reshape(-1)))
        assert raw[16:].data_ptr() % 32 == 16
        return raw[16:]

    o_S, o_sw = shifted(S0), shifted(sw0)
    solver06.adapt_inputs(shifted(com), shifted(js), o_S, o_sw, n, stream=st)
    o_cmd = shifted(np.zeros(n, dtype=TORQUE_CMD_DTYPE))
    solver06.torque_cmd(o_S, shifted(out), o_cmd, n, stream=st)
    torch.cuda.synchronize()
    assert _host(o_S, STATE_DTYPE).tobytes() == got_S.tobytes() and _host(o_sw, SWING_DTYPE).tobytes() == got_sw.tobytes()
    assert _host(o_cmd, TORQUE_CMD_DTYPE).tobytes() == cmd.tobytes()
    m = 19
    t_S, t_sw = _dev(S0[:m]), _dev(sw0[:m])
    solver06.adapt_inputs(_dev(com[:m]), _dev(js[:m]), t_S, t_sw, m, stream=st)
    t_cmd = torch.zeros(m * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    solver06.torque_cmd(t_S, _dev(out[:m]), t_cmd, m, stream=st)
    torch.cuda.synchronize()
    assert _host(t_S, STATE_DTYPE).tobytes() == got_S[:m].tobytes() and _host(t_sw, SWING_DTYPE).tobytes() == got_sw[:m].tobytes()
    assert _host(t_cmd, TORQUE_CMD_DTYPE).tobytes() == cmd[:m].tobytes()


@pytest.mark.gpu
def test_whole_tick_from_messages_gpu(solver06, params06):
    """messages -> adapt -> plan -> tick -> torque command on one stream, against the same chain of oracle calls."""
    import oracle
    import torch

    n = 4096
    pp, gains = default_plan_params(), default_joint_gains()
    S = states.generate_states(n, 41, masks="mixed", profile="light")  # slow robots: planned footholds stay within reach
    com, js = make_msgs(n, 42)
    js["position"] = S["q"].reshape(n, 4, 3).transpose(0, 2, 1).reshape(n, 12)  # a reachable posture, in message order
    rot = S["Rwb"].reshape(n, 3, 3)
    from scipy.spatial.transform import Rotation

    com["orientation"] = Rotation.from_matrix(rot).as_quat()
    com["position"] = S["x"]
    com["linear"] = S["xdot"]
    com["angular"] = S["w"]
    plan = make_plan(S, 43, all_replan=True)
    plan["phase"] = np.random.default_rng(44).uniform(0.83, 1.0, size=(n, 4))
    sw = np.zeros(n, dtype=SWING_DTYPE)
    d_S, d_plan, d_sw = _dev(S), _dev(plan), _dev(sw)
    d_out = torch.zeros(n * OUT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    d_cmd = torch.zeros(n * TORQUE_CMD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    solver06.adapt_inputs(_dev(com), _dev(js), d_S, d_sw, n, stream=st)
    solver06.plan(d_S, d_plan, d_sw, n, stream=st)
    solver06.tick_packed(d_S, d_sw, d_out, n, stream=st)
    solver06.torque_cmd(d_S, d_out, d_cmd, n, stream=st)
    torch.cuda.synchronize()
    S_ref, plan_ref, sw_ref = S.copy(), plan.copy(), sw.copy()
    oracle.adapt_inputs(params06, com, js, S_ref, sw_ref)
    oracle.plan_batch(pp, S_ref, plan_ref, sw_ref)
    out_ref = oracle.tick_batch(params06, gains, S_ref, sw_ref, os.cpu_count() or 1)
    out = _host(d_out, OUT_DTYPE)
    assert np.array_equal(out["status"], out_ref["status"])
    assert rel_err(out["grf_body"], out_ref["grf_body"]) <= 1e-5
    # planned footholds of fast-moving robots can lie out of the leg's reach: legInverseKinematics then returns NaN in the
    # reference too (acos of |arg| > 1, kinematics.cpp:117-160); the NaN pattern must agree and the rest must match
    unreachable = np.isnan(out_ref["tau"])
    assert np.array_equal(np.isnan(out["tau"]), unreachable) and unreachable.mean() < 0.5
    # ... and at the edge of reach the leg Jacobian is singular, where inv/pinv return unpinned garbage of any size on both
    # sides: compare the well-conditioned entries
    sane = ~unreachable & (np.abs(out_ref["tau"]) < 1e3)
    assert sane.mean() > 0.6
    bad = np.abs(out["tau"] - out_ref["tau"])[sane] > 1e-5 * np.maximum(np.abs(out_ref["tau"][sane]), 1.0)
    assert bad.mean() < 1e-3, (bad.mean(), sane.mean(), unreachable.mean())  # a leg at the very edge of reach may still slip through
    cmd = _host(d_cmd, TORQUE_CMD_DTYPE)
    assert (cmd["count"][out["status"] == 0] == 12).all() and np.nanmax(np.abs(cmd["torque"])) <= 20.0
