"""CPU tests of the oracle itself: pinned against the reference's stored kinematics outputs,
certified by the solver-free KKT checker, cross-checked by brute-force active-set enumeration."""
import itertools
import json
import os

import numpy as np
import pytest

import oracle
from oracle.kkt import certificate
from quadruped_control_b200 import STATE_DTYPE, default_params, states

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def nb():
    with open(os.path.join(GOLD, "notebook_kinematics.json")) as f:
        return json.load(f)


def test_params_match_reference_config(params08):
    # mit_cheetah_config.yaml:66-99, kinematics.cpp:23-47
    po = oracle.default_params()
    assert bytes(po) == bytes(params08)
    assert po.mu == 0.8 and po.mass == 11.0 and po.fzmin == 10.0 and po.fzmax == 120.0
    assert np.allclose(np.array(po.S[:]).reshape(6, 6), np.diag([1, 1, 1, 10, 10, 5]))
    assert np.allclose(np.array(po.W[:]).reshape(12, 12), 1e-5 * np.eye(12))
    assert po.max_iter == 200 and po.clamp_tau == 0


def test_forward_kinematics_matches_notebook(nb, params08):
    q = nb["q"]
    assert np.allclose(oracle.forward_kinematics(params08, 0, q), nb["foot_RL"], atol=5e-9)
    assert np.allclose(oracle.forward_kinematics(params08, 2, q), nb["foot_RR"], atol=5e-9)
    # vectorised host generator uses the same formula
    q12 = np.tile(q, 4)
    feet = states.forward_kinematics(q12, params08).reshape(4, 3)
    assert np.allclose(feet[0], nb["foot_RL"], atol=5e-9) and np.allclose(feet[2], nb["foot_RR"], atol=5e-9)
    for leg in range(4):
        assert np.allclose(feet[leg], oracle.forward_kinematics(params08, leg, q), atol=1e-15)


def test_jacobian_matches_notebook(nb, params08):
    q = nb["q"]
    for leg, key in ((0, "J_left"), (1, "J_left"), (2, "J_right"), (3, "J_right")):
        assert np.allclose(oracle.leg_jacobian(params08, leg, q), nb[key], atol=5e-9)


def test_jacobian_is_derivative_of_fk(params08):
    rng = np.random.default_rng(0)
    for leg in range(4):
        q = rng.uniform(-1, 1, 3) + np.array([0, 0.9, -1.9])
        J = oracle.leg_jacobian(params08, leg, q)
        h = 1e-6
        for k in range(3):
            dq = np.zeros(3)
            dq[k] = h
            num = (oracle.forward_kinematics(params08, leg, q + dq) - oracle.forward_kinematics(params08, leg, q - dq)) / (2 * h)
            assert np.allclose(J[:, k], num, atol=1e-8)


def test_stance_pose_feet_on_ground(nb, params08):
    s = states.stance_state(params08)
    feet = s["feet"][0].reshape(4, 3)
    assert np.allclose(feet[0], nb["stance_foot_RL"], atol=1e-9)
    assert np.allclose(feet[1], nb["stance_foot_FL"], atol=1e-9)
    assert np.allclose(feet[:, 2], -nb["stance_com_height"], atol=2e-5)


def test_angle_axis_matches_scipy_and_pi_branch():
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(1)
    for _ in range(200):
        rv = rng.normal(size=3)
        rv *= rng.uniform(0, np.pi - 1e-3) / np.linalg.norm(rv)
        R = Rotation.from_rotvec(rv).as_matrix()
        assert np.allclose(oracle.angle_axis_total(R), rv, atol=1e-10)
    assert np.allclose(oracle.angle_axis_total(np.eye(3)), 0.0)
    # trace <= 0 branches: rotations by ~pi about each axis
    for axis in range(3):
        rv = np.zeros(3)
        rv[axis] = np.pi - 1e-4
        R = Rotation.from_rotvec(rv).as_matrix()
        assert np.allclose(oracle.angle_axis_total(R), rv, atol=1e-9)


def test_config1_known_answer(params08):
    # SURVEY.md App. C.3 (numpy FP64, KKT-verified; regression value, not qpOASES output)
    out, fw = oracle.control(params08, states.stance_state(params08))
    assert out["status"][0] == 0
    assert np.allclose(fw[2::3], [25.98995628, 18.48567936, 25.98995628, 18.48567936], atol=5e-8)
    assert np.allclose(out["grf_body"][0], -fw, atol=1e-12)  # R = I
    tau = out["tau"][0].reshape(4, 3)
    assert np.allclose(tau[0], [-2.3582535, 0.8581635, 5.1471025], atol=5e-7)
    assert np.allclose(tau[1], [-1.6773371, 0.6103786, 3.6609402], atol=5e-7)
    assert np.allclose(tau[2], [2.3582535, 0.8581635, 5.1471025], atol=5e-7)
    assert np.allclose(tau[3], [1.6773371, 0.6103786, 3.6609402], atol=5e-7)
    qp = oracle.assemble(params08, states.stance_state(params08))
    assert np.allclose(qp["b"], [0, 0, 88.9515, 0, 0, 0], atol=1e-9)
    ev = np.linalg.eigvalsh(qp["Q"])
    assert abs(ev[0] - 2e-5) < 1e-9 and abs(ev[-1] - 14.142) < 1e-2


def test_assembly_matches_dense_numpy_restatement(params06):
    """Second, independent (numpy) statement of balance_controller.cpp:126-153, 237-330."""
    from scipy.spatial.transform import Rotation

    p = params06
    S = np.array(p.S[:]).reshape(6, 6)
    W = np.array(p.W[:]).reshape(12, 12)
    Ib = np.array(p.Ib[:]).reshape(3, 3)
    for s in states.generate_states(50, 99, masks="mixed"):
        qp = oracle.assemble(p, np.array([s]))
        R, Rd = s["Rwb"].reshape(3, 3), s["Rwb_d"].reshape(3, 3)
        a = 100.0 * (s["x_d"] - s["x"]) + 50.0 * (s["xdot_d"] - s["xdot"])
        a[2] += 0.15 * 11.0 * 9.81
        e = Rotation.from_matrix(Rd @ R.T).as_rotvec()
        al = 5000.0 * e + 500.0 * (s["w_d"] - s["w"])
        A = np.zeros((6, 12))
        for i in range(4):
            r = R @ s["feet"][3 * i:3 * i + 3]
            A[:3, 3 * i:3 * i + 3] = np.eye(3)
            A[3:, 3 * i:3 * i + 3] = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        Iw = R @ Ib @ R.T
        b = np.concatenate([11.0 * (a + [0, 0, -9.81]), Iw @ al + np.cross(s["w_d"], Iw @ s["w_d"])])
        assert np.allclose(qp["A"], A, atol=1e-14)
        assert np.allclose(qp["b"], b, rtol=1e-9, atol=1e-9)
        assert np.allclose(qp["Q"], 2 * (A.T @ S @ A + W), rtol=1e-12, atol=1e-13)
        assert np.allclose(qp["c"], -2 * A.T @ S @ b, rtol=1e-9, atol=1e-8)
        for leg in range(4):
            rows = slice(5 * leg, 5 * leg + 5)
            if s["contact"][leg]:
                assert np.array_equal(qp["lb"][rows], [-1e6, -1e6, 0, 0, 10.0])
                assert np.array_equal(qp["ub"][rows], [0, 0, 1e6, 1e6, 120.0])
            else:
                assert not qp["lb"][rows].any() and not qp["ub"][rows].any()
        Cf = np.array([[1, 0, -0.6], [0, 1, -0.6], [0, 1, 0.6], [1, 0, 0.6], [0, 0, 1]])
        C = np.zeros((20, 12))
        for leg in range(4):
            C[5 * leg:5 * leg + 5, 3 * leg:3 * leg + 3] = Cf
        assert np.array_equal(qp["C"], C)


def test_kff_quirk_index1_twice():
    # balance_controller.cpp:139 adds kff[5]*w_d[2] to wdot_d[1]
    p = default_params(0.6)
    p.kff[5] = 2.0
    s = states.stance_state(p)
    s["w_d"][0] = (0.0, 0.0, 0.5)
    b1 = oracle.assemble(p, s)["b"]
    p.kff[5] = 0.0
    b0 = oracle.assemble(p, s)["b"]
    Iw = np.array(p.Ib[:]).reshape(3, 3)
    assert np.allclose(b1[3:] - b0[3:], Iw @ [0.0, 2.0 * 0.5, 0.0], atol=1e-12)


@pytest.mark.parametrize("profile", ["default", "light", "stress"])
@pytest.mark.parametrize("masks", ["all4", "mixed"])
def test_oracle_answers_carry_kkt_certificate(params06, profile, masks):
    S = states.generate_states(150, 4242, profile=profile, masks=masks)
    for i in range(len(S)):
        qp = oracle.assemble(params06, S[i:i + 1])
        st, x, lam, it = oracle.qp_solve(qp["Q"], qp["c"], qp["C"], qp["lb"], qp["ub"])
        assert st == 0
        cert = certificate(qp["Q"], qp["c"], qp["C"], qp["lb"], qp["ub"], x)
        assert cert["infeas"] <= 1e-9
        assert cert["stat_rel"] <= 1e-11, cert
        assert cert["dist_bound"] <= 1e-4  # ||x - x*|| in N; typical forces are 10..100 N


def _brute_force(Q, c, N, b):
    """Enumerate active sets of  n_j'x >= b_j : the KKT point with all multipliers >= 0 that is feasible."""
    n, m = Q.shape[0], N.shape[0]
    best = None
    for k in range(0, min(n, m) + 1):
        for act in itertools.combinations(range(m), k):
            Na = N[list(act)]
            if k and np.linalg.matrix_rank(Na) < k:
                continue
            K = np.block([[Q, -Na.T], [Na, np.zeros((k, k))]])
            sol = np.linalg.solve(K, np.concatenate([-c, b[list(act)]]))
            x, lam = sol[:n], sol[n:]
            if (lam >= -1e-9).all() and (N @ x - b >= -1e-8).all():
                val = 0.5 * x @ Q @ x + c @ x
                if best is None or val < best[0]:
                    best = (val, x)
    return best[1]


def test_qp_solver_against_active_set_enumeration():
    """Two stance legs (6 variables, 12 one-sided rows): every active set can be enumerated."""
    p = default_params(0.6)
    S = states.generate_states(6, 31337)
    S["contact"] = [1, 0, 0, 1]
    for i in range(len(S)):
        qp = oracle.assemble(p, S[i:i + 1])
        st, x, lam, it = oracle.qp_solve(qp["Q"], qp["c"], qp["C"], qp["lb"], qp["ub"])
        assert st == 0 and np.allclose(x[3:9], 0.0, atol=1e-9)
        idx = [0, 1, 2, 9, 10, 11]
        Qs, cs = qp["Q"][np.ix_(idx, idx)], qp["c"][idx]
        rows = list(range(0, 5)) + list(range(15, 20))
        Cs = qp["C"][np.ix_(rows, idx)]
        N = np.vstack([Cs, -Cs])
        b = np.concatenate([qp["lb"][rows], -qp["ub"][rows]])
        keep = np.abs(b) < 1e5  # the +-1e6 sides can never be active
        xb = _brute_force(Qs, cs, N[keep], b[keep])
        assert np.allclose(x[idx], xb, rtol=1e-7, atol=1e-7)


def test_swing_rows_equal_variable_elimination(params06):
    """The reference encodes a swing leg as five zero-equality rows (balance_controller.cpp:312-316);
    that must equal deleting the leg's variables."""
    S = states.generate_states(20, 555)
    S["contact"][:, 1] = 0
    for i in range(len(S)):
        qp = oracle.assemble(params06, S[i:i + 1])
        st, x, lam, it = oracle.qp_solve(qp["Q"], qp["c"], qp["C"], qp["lb"], qp["ub"])
        idx = [0, 1, 2, 6, 7, 8, 9, 10, 11]
        rows = [r for r in range(20) if not 5 <= r < 10]
        st2, x2, _, _ = oracle.qp_solve(qp["Q"][np.ix_(idx, idx)], qp["c"][idx], qp["C"][np.ix_(rows, idx)],
                                        qp["lb"][rows], qp["ub"][rows])
        assert st == 0 and st2 == 0
        assert np.allclose(x[3:6], 0.0, atol=1e-9) and np.allclose(x[idx], x2, rtol=1e-8, atol=1e-8)


def test_failure_modes():
    p = default_params(0.6)
    s = states.stance_state(p)
    s["x"][0, 0] = np.nan
    out, _ = oracle.control(p, s)
    assert out["status"][0] == 2 and not out["grf_body"].any() and not out["tau"].any()
    p.max_iter = 1
    S = states.generate_states(64, 1, profile="stress")
    O = oracle.control_batch(p, S)
    assert (O["status"] == 1).any()
    bad = O["status"] != 0
    assert not O["grf_body"][bad].any() and not O["tau"][bad].any()
    # infeasible rows are reported, not looped on
    st, x, lam, it = oracle.qp_solve(np.eye(2), np.zeros(2), np.array([[1.0, 0.0], [1.0, 0.0]]), [1.0, -5.0], [2.0, 0.0])
    assert st == 2


def test_golden_fixture_is_current(params06, params08):
    g = np.load(os.path.join(GOLD, "balance_golden.npz"))
    S = g["states"].reshape(-1).view(STATE_DTYPE)
    for mu, p in ((0.6, params06), (0.8, params08)):
        O = oracle.control_batch(p, S, 2)
        assert np.array_equal(O["status"], g[f"status_mu{mu}"])
        assert np.allclose(O["grf_body"], g[f"grf_mu{mu}"], rtol=1e-9, atol=1e-9)
        assert np.allclose(O["tau"], g[f"tau_mu{mu}"], rtol=1e-9, atol=1e-9)


def test_batch_threads_agree(params06):
    S = states.generate_states(501, 77, masks="mixed")
    a = oracle.control_batch(params06, S, 1)
    b = oracle.control_batch(params06, S, 5)
    assert a.tobytes() == b.tobytes()


def test_size_independent_properties_of_the_solution(params06):
    """Properties the domain offers, used at full size by the GPU tests too: (1) scaling S and W by a common factor scales
    the objective and leaves the minimiser alone; (2) mirroring the robot left <-> right (y -> -y) mirrors the forces;
    (3) inactive rows do not matter: tightening fzmax to just above the largest returned fz changes nothing;
    (4) body-frame results are invariant under a quarter turn of the world about z."""
    S = states.generate_states(256, 515, masks="mixed")
    ref = oracle.control_batch(params06, S)
    assert (ref["status"] == 0).all()
    # (1)
    p2 = params06.copy()
    for i in range(36):
        p2.S[i] = 3.0 * params06.S[i]
    for i in range(144):
        p2.W[i] = 3.0 * params06.W[i]
    out2 = oracle.control_batch(p2, S)
    assert np.abs(out2["grf_body"] - ref["grf_body"]).max() <= 1e-7 * np.abs(ref["grf_body"]).max()
    # (2) reflection M = diag(1, -1, 1): vectors v -> M v, rotations R -> M R M, pseudo-vectors (w) -> -M w; legs RL<->RR, FL<->FR
    M = np.diag([1.0, -1.0, 1.0])
    perm = [2, 3, 0, 1]
    Sm = S.copy()
    for f in ("Rwb", "Rwb_d"):
        Sm[f] = (M @ S[f].reshape(-1, 3, 3) @ M).reshape(-1, 9)
    for f in ("x", "xdot", "x_d", "xdot_d"):
        Sm[f] = S[f] @ M
    for f in ("w", "w_d"):
        Sm[f] = -(S[f] @ M)
    Sm["feet"] = (S["feet"].reshape(-1, 4, 3)[:, perm] @ M).reshape(-1, 12)
    Sm["contact"] = S["contact"][:, perm]
    qm = S["q"].reshape(-1, 4, 3)[:, perm].copy()
    qm[..., 0] *= -1.0  # hip abduction changes sign under the mirror; thigh and calf angles do not
    Sm["q"] = qm.reshape(-1, 12)
    outm = oracle.control_batch(params06, Sm)
    assert (outm["status"] == 0).all()
    want = (ref["grf_body"].reshape(-1, 4, 3)[:, perm] @ M).reshape(-1, 12)
    assert np.abs(outm["grf_body"] - want).max() <= 1e-6 * np.abs(want).max()
    wt = ref["tau"].reshape(-1, 4, 3)[:, perm].copy()
    wt[..., 0] *= -1.0
    assert np.abs(outm["tau"] - wt.reshape(-1, 12)).max() <= 1e-6 * np.abs(wt).max()
    # (4) a quarter turn of the world about z maps the friction pyramids onto themselves and S, the gains and gravity are
    #     symmetric in x and y: body-frame forces and torques must not change
    G = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    Sg = S.copy()
    for f in ("Rwb", "Rwb_d"):
        Sg[f] = (G @ S[f].reshape(-1, 3, 3)).reshape(-1, 9)
    for f in ("x", "xdot", "w", "x_d", "xdot_d", "w_d"):
        Sg[f] = S[f] @ G.T
    outg = oracle.control_batch(params06, Sg)
    assert (outg["status"] == 0).all()
    assert np.abs(outg["grf_body"] - ref["grf_body"]).max() <= 1e-6 * np.abs(ref["grf_body"]).max()
    assert np.abs(outg["tau"] - ref["tau"]).max() <= 1e-6 * np.abs(ref["tau"]).max()
    # (3)
    fz_world = -np.einsum("nij,nlj->nli", S["Rwb"].reshape(-1, 3, 3), ref["grf_body"].reshape(-1, 4, 3))[..., 2]
    inner = fz_world.max(axis=1) < params06.fzmax - 1.0  # QPs whose fzmax rows are all inactive
    assert inner.sum() > 20
    p3 = params06.copy()
    p3.fzmax = float(fz_world[inner].max()) + 1e-3
    out3 = oracle.control_batch(p3, np.ascontiguousarray(S[inner]))
    assert np.abs(out3["grf_body"] - ref["grf_body"][inner]).max() <= 1e-7 * np.abs(ref["grf_body"]).max()
