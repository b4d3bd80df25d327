"""oracle/_ref: the reference's OWN sources (balance_controller.cpp, kinematics.cpp, gait.cpp,
math/numerics.cpp, compiled from /root/reference against stand-in Armadillo/qpOASES/ROS/rigid3d headers)
pin the oracle's restatement of everything except the third-party QP solve and log map."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import rel_err
from quadruped_control_b200 import default_params, states

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def reference_library():
    """oracle/_ref/libqpb_ref.so must exist: built here from /root/reference (oracle/ref_build.sh) or shipped prebuilt
    with the tree (it is git-ignored, not gpurun-ignored).  Its absence is a failure, not a skip: without it the oracle's
    restatement of the reference's own lines is unpinned."""
    if not oracle.ref_available():
        pytest.fail("oracle/_ref/libqpb_ref.so is missing and /root/reference is absent: run oracle/ref_build.sh where the "
                    "reference lives and ship oracle/_ref/ with the tree")


def test_reference_kinematics_reproduce_notebook_vectors():
    nb = json.load(open(os.path.join(GOLD, "notebook_kinematics.json")))
    feet = oracle.ref_forward_kinematics(np.tile(nb["q"], 4)).reshape(4, 3)
    assert np.allclose(feet[0], nb["foot_RL"], atol=5e-9) and np.allclose(feet[2], nb["foot_RR"], atol=5e-9)
    for leg, key in ((0, "J_left"), (1, "J_left"), (2, "J_right"), (3, "J_right")):
        assert np.allclose(oracle.ref_leg_jacobian(leg, nb["q"]), nb[key], atol=5e-9)


def test_oracle_kinematics_equal_reference_sources(params08):
    rng = np.random.default_rng(3)
    for _ in range(50):
        q = states.STANCE_Q + rng.uniform(-0.8, 0.8, 12)
        feet = oracle.ref_forward_kinematics(q)
        for leg in range(4):
            ql = q[3 * leg:3 * leg + 3]
            assert np.array_equal(oracle.forward_kinematics(params08, leg, ql), feet[3 * leg:3 * leg + 3])
            assert np.array_equal(oracle.leg_jacobian(params08, leg, ql), oracle.ref_leg_jacobian(leg, ql))


@pytest.mark.parametrize("profile", ["default", "light", "stress"])
@pytest.mark.parametrize("masks", ["all4", "mixed"])
def test_oracle_equals_reference_sources(params06, profile, masks):
    S = states.generate_states(1500, 8080, profile=profile, masks=masks)
    ref = oracle.ref_control_batch(params06, S, 2)
    orc = oracle.control_batch(params06, S, 2)
    assert np.array_equal(ref["status"], orc["status"]) and (ref["status"] == 0).all()
    assert rel_err(orc["grf_body"], ref["grf_body"]) <= 1e-9
    assert rel_err(orc["tau"], ref["tau"]) <= 1e-9
    # legs in swing are absent from the reference's maps (bc.cpp:223-228) -> zero in the flat record
    swing = np.repeat(S["contact"] == 0, 3, axis=1)
    assert not ref["grf_body"][swing].any() and not ref["tau"][swing].any()


def test_reference_sources_general_parameters_and_quirks():
    rng = np.random.default_rng(9)
    p = default_params(0.5)
    A6 = rng.normal(size=(6, 6))
    A12 = rng.normal(size=(12, 12))
    p.S[:] = (np.diag([1, 1, 1, 10, 10, 5.0]) + 0.05 * A6 @ A6.T).ravel().tolist()
    p.W[:] = (1e-5 * np.eye(12) + 2e-6 * A12 @ A12.T).ravel().tolist()
    p.kff[:] = [0.3, -0.2, 0.15, 0.5, -0.4, 0.7]  # exercises the kff[5] -> index 1 quirk (bc.cpp:139)
    p.fzmin, p.fzmax = 0.0, 90.0
    S = states.generate_states(800, 4, profile="stress", masks="mixed")
    ref = oracle.ref_control_batch(p, S, 2)
    orc = oracle.control_batch(p, S, 2)
    assert np.array_equal(ref["status"], orc["status"])
    assert rel_err(orc["grf_body"], ref["grf_body"]) <= 1e-8 and rel_err(orc["tau"], ref["tau"]) <= 1e-8


def test_golden_fixture_matches_reference_sources(params06, params08):
    from quadruped_control_b200 import STATE_DTYPE

    g = np.load(os.path.join(GOLD, "balance_golden.npz"))
    S = np.ascontiguousarray(g["states"]).reshape(-1).view(STATE_DTYPE)
    for mu, p in ((0.6, params06), (0.8, params08)):
        ref = oracle.ref_control_batch(p, S)
        assert np.array_equal(ref["status"], g[f"status_mu{mu}"])
        assert rel_err(ref["grf_body"], g[f"grf_mu{mu}"]) <= 1e-8
