"""Swing-leg half of the control tick (commander_node.cpp:482-533): oracle pinned against the
reference's own sources on the CPU, CUDA path against the oracle on the GPU."""
import numpy as np
import pytest

import oracle
from conftest import rel_err
from quadruped_control_b200 import OUT_DTYPE, default_joint_gains, default_params, states

NCPU = 4


def _data(n, seed, masks="mixed", profile="default"):
    S = states.generate_states(n, seed, profile=profile, masks=masks)
    return S, states.generate_swing(S, seed + 1)


def test_joint_gains_defaults_match_config():
    g = oracle.default_joint_gains()  # mit_cheetah_config.yaml:50-53
    assert bytes(g) == bytes(default_joint_gains())
    assert list(g.kff) == [0.0, 0.0, 0.0] and list(g.kp) == [40.0, 40.0, 50.0] and list(g.kd) == [1.0, 1.0, 1.0]


def test_inverse_kinematics_round_trip_and_jacobian_inverse(params08):
    rng = np.random.default_rng(4)
    for _ in range(100):
        for leg in range(4):
            q = states.STANCE_Q[3 * leg:3 * leg + 3] + rng.uniform(-0.4, 0.4, 3)
            foot = oracle.forward_kinematics(params08, leg, q)
            assert np.allclose(oracle.leg_inverse_kinematics(params08, leg, foot), q, atol=1e-9)
            Ji, kind = oracle.leg_jacobian_inverse(params08, leg, q)
            assert kind == 0 and np.allclose(Ji @ oracle.leg_jacobian(params08, leg, q), np.eye(3), atol=1e-9)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not available")
def test_swing_kinematics_equal_reference_sources(params08):
    rng = np.random.default_rng(5)
    for _ in range(50):
        for leg in range(4):
            q = states.STANCE_Q[3 * leg:3 * leg + 3] + rng.uniform(-0.4, 0.4, 3)
            foot = oracle.forward_kinematics(params08, leg, q) + rng.normal(0, 0.02, 3)
            assert np.array_equal(oracle.leg_inverse_kinematics(params08, leg, foot), oracle.ref_leg_inverse_kinematics(leg, foot))
            assert np.allclose(oracle.leg_jacobian_inverse(params08, leg, q)[0], oracle.ref_leg_jacobian_inverse(leg, q), rtol=1e-12, atol=1e-12)
    # out-of-reach target: d is clamped to 1 (kinematics.cpp:133-136), both sides agree on the stretched leg
    far = np.array([-0.196, 0.05 + 0.077, -0.9])
    assert np.allclose(oracle.leg_inverse_kinematics(params08, 0, far), oracle.ref_leg_inverse_kinematics(0, far), atol=1e-12)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not available")
@pytest.mark.parametrize("clamp", [0, 1])
def test_tick_oracle_equals_reference_sources(clamp):
    p = default_params(0.6)
    p.clamp_tau = clamp
    g = default_joint_gains()
    g.kff[:] = [0.3, -0.2, 0.1]
    S, SW = _data(1500, 31)
    ref = oracle.ref_tick_batch(p, g, S, SW, 2)
    orc = oracle.tick_batch(p, g, S, SW, 2)
    assert np.array_equal(ref["status"], orc["status"])
    assert rel_err(orc["grf_body"], ref["grf_body"]) <= 1e-9 and rel_err(orc["tau"], ref["tau"]) <= 1e-9
    sw = np.repeat(S["contact"] == 0, 3, axis=1)
    assert np.abs(orc["tau"][sw]).min() > 0.0  # every swing joint gets a PD torque
    if clamp:
        assert orc["tau"].max() <= 20.0 and orc["tau"].min() >= -20.0 and (np.abs(orc["tau"]) == 20.0).any()
    # stance part of the tick equals the balance-only call
    bal = oracle.control_batch(p, S, 2)
    st = ~sw
    assert np.array_equal(orc["grf_body"], bal["grf_body"]) and np.array_equal(orc["tau"][st], bal["tau"][st])


def test_angle_wrapping_in_joint_pd():
    """joint_controller.cpp:27-31 wraps both angles to [0, 2pi) and the error to [-pi, pi)."""
    p, g = default_params(0.6), default_joint_gains()
    S, SW = _data(64, 9)
    S2 = S.copy()
    S2["q"] += 2 * np.pi * np.random.default_rng(0).integers(-2, 3, size=(64, 12))
    S2["feet"] = S["feet"]
    a = oracle.tick_batch(p, g, S, SW)
    b = oracle.tick_batch(p, g, S2, SW)
    sw = np.repeat(S["contact"] == 0, 3, axis=1)
    assert np.allclose(a["tau"][sw], b["tau"][sw], atol=1e-9)


# ---------------------------------------------------------------- GPU ---------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("clamp", [0, 1])
def test_gpu_tick_parity(built, clamp):
    import torch

    from quadruped_control_b200 import lib

    p = default_params(0.6)
    p.clamp_tau = clamp
    g = default_joint_gains()
    g.kff[:] = [0.3, -0.2, 0.1]
    S, SW = _data(20001, 77, profile="stress")
    solver = lib.BalanceSolver(p)
    solver.set_joint_gains(g)
    out = solver.tick_host(S, SW)
    ref = oracle.tick_batch(p, g, S, SW, NCPU)
    assert np.array_equal(out["status"], ref["status"])
    assert rel_err(out["grf_body"], ref["grf_body"]) <= 1e-5
    assert rel_err(out["tau"], ref["tau"]) <= 1e-5
    assert rel_err(out["tau"], ref["tau"]) <= 1e-7  # measured ~1e-9
    # device-resident entry point gives the same bytes
    d_s = torch.from_numpy(S.view(np.uint8).reshape(-1).copy()).cuda()
    d_w = torch.from_numpy(SW.view(np.uint8).reshape(-1).copy()).cuda()
    d_o = torch.zeros(len(S) * 256, dtype=torch.uint8, device="cuda")
    solver.tick_packed(d_s, d_w, d_o, len(S))
    torch.cuda.synchronize()
    assert d_o.cpu().numpy().view(OUT_DTYPE).tobytes() == out.tobytes()
    # the balance-only call is untouched by the joint gains
    bal = solver.control_host(S)
    st = np.repeat(S["contact"] != 0, 3, axis=1)
    assert np.array_equal(bal["grf_body"], out["grf_body"]) and np.array_equal(bal["tau"][st], out["tau"][st])
    solver.close()


@pytest.mark.gpu
def test_gpu_tick_failure_paths(built):
    from quadruped_control_b200 import lib

    p = default_params(0.6)
    p.max_iter = 2
    g = default_joint_gains()
    S, SW = _data(512, 5, profile="stress")
    S["x"][7, 1] = np.inf
    solver = lib.BalanceSolver(p)
    out = solver.tick_host(S, SW)
    ref = oracle.tick_batch(p, g, S, SW, NCPU)
    assert np.array_equal(out["status"] == 2, ref["status"] == 2) and out["status"][7] == 2 and (out["status"] == 1).any()
    assert not out["tau"][7].any() and not out["grf_body"][7].any()  # broken state: nothing is commanded
    over = out["status"] == 1  # QP gave up: the reference still publishes the swing torques alone
    assert not out["grf_body"][over].any()
    # two different solvers need not hit a 2-iteration limit on exactly the same QPs: compare where they agree
    same = out["status"] == ref["status"]
    assert same.mean() > 0.95 and rel_err(out["tau"][same], ref["tau"][same]) <= 1e-7
    sw = np.repeat(S["contact"] == 0, 3, axis=1) & over[:, None]
    assert np.abs(out["tau"][sw]).min() > 0.0
    solver.close()
