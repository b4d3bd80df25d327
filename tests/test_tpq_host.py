"""CPU checks of the thread-per-QP solver core (quadruped_control_b200/csrc/qpb_tpq_core.h).

The CUDA kernel balance_qp_tpq_kernel runs this very source, one QP per thread; tests/host_tpq/tpq_host.cpp compiles
it with g++ so the algorithm -- range-space Goldfarb-Idnani, face family, warm start -- is checked against the oracle
here, without a GPU.  (The product path never loads this library; the GPU tests check the kernel itself.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import rel_err
from quadruped_control_b200 import OUT_DTYPE, default_params, states

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_tpq", "tpq_host.cpp")
LIB = os.path.join(HERE, "host_tpq", "libtpq_host.so")
NCPU = os.cpu_count() or 1


@pytest.fixture(scope="module")
def host():
    csrc = os.path.join(os.path.dirname(HERE), "quadruped_control_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("qpb_tpq_core.h", "qpb_stages.h", "qpb_math.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    L = ctypes.CDLL(LIB)
    L.tpq_host_control_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]

    def run(params, S, lpq=2, polish=1):
        S = np.ascontiguousarray(S)
        out = np.zeros(len(S), dtype=OUT_DTYPE)
        rc = L.tpq_host_control_batch(ctypes.byref(params), S.ctypes.data, len(S), out.ctypes.data, lpq, polish)
        return rc, out

    return run


def _check(run, params, S, tol, lpq=2):
    rc, out = run(params, S, lpq)
    assert rc == 0
    ref = oracle.control_batch(params, S, NCPU)
    assert np.array_equal(out["status"], ref["status"])
    ef, et = rel_err(out["grf_body"], ref["grf_body"]), rel_err(out["tau"], ref["tau"])
    assert ef <= tol and et <= tol, (ef, et)
    return out, ref


@pytest.mark.parametrize("profile", ["default", "light", "stress"])
@pytest.mark.parametrize("masks", ["all4", "mixed"])
def test_core_matches_oracle(host, params06, profile, masks):
    S = states.generate_states(6000, 20260102 if masks == "all4" else 20260103, profile=profile, masks=masks)
    out, ref = _check(host, params06, S, 1e-7)
    assert out["iters"].max() < 64


def test_lanes_per_qp_are_the_same_algorithm(host, params06):
    """One, two or four lanes per QP (the kernel's template parameter) only distribute the legs: working sets,
    iteration counts and results must be identical."""
    S = states.generate_states(3000, 20260103, profile="stress", masks="mixed")
    outs = [host(params06, S, lpq)[1] for lpq in (1, 2, 4)]
    for o in outs[1:]:
        assert o.tobytes() == outs[0].tobytes()
    _check(host, params06, S, 1e-7, lpq=4)
    _check(host, params06, S, 1e-7, lpq=1)


def test_core_config1_and_every_contact_mask(host, params06, params08):
    out, _ = _check(host, params08, states.stance_state(params08), 1e-10)
    assert out["iters"][0] == 0
    S = states.generate_states(4096, 606, profile="stress", masks="mixed")
    S["contact"] = (np.arange(len(S))[:, None] % 16 >> np.arange(4)) & 1
    out, _ = _check(host, params06, S, 1e-7)
    swing = np.repeat(S["contact"] == 0, 3, axis=1)
    assert not out["grf_body"][swing].any() and not out["tau"][swing].any()


def test_core_degenerate_parameter_regimes(host):
    """The regimes of tests/test_gpu_parity.py::test_degenerate_and_extreme_parameter_regimes: the face family must
    cover them without ever leaving it (no QP is handed back)."""
    rng = np.random.default_rng(12)
    S = states.generate_states(3000, 91, profile="stress", masks="mixed")
    S["x_d"][:, 2] = S["x"][:, 2] - rng.uniform(0.0, 0.6, len(S))
    S2 = states.generate_states(3000, 92, profile="default", masks="mixed")
    cases = []
    p = default_params(0.6); p.fzmin = 0.0; cases.append((p, S))
    p = default_params(0.6); p.fzmin = 35.0; p.fzmax = 35.0; cases.append((p, S))
    p = default_params(0.01); cases.append((p, S))
    p = default_params(5.0); cases.append((p, S))
    p = default_params(0.6); p.mass = 60.0; cases.append((p, S2))
    p = default_params(0.6); p.W[:] = (1e-1 * np.eye(12)).ravel().tolist(); cases.append((p, S2))
    for p, Sin in cases:
        _check(host, p, Sin, 1e-6)


def test_core_general_S_and_gains(host):
    """Dense SPD S, non-diagonal inertia, all feed-forward gains, torque clamp: still the fast path (only W must be w I)."""
    rng = np.random.default_rng(5)
    p = default_params(0.45)
    A6 = rng.normal(size=(6, 6))
    p.S[:] = (np.diag([1, 1, 1, 10, 10, 5.0]) + 0.05 * (A6 @ A6.T)).ravel().tolist()
    A3 = rng.normal(size=(3, 3))
    p.Ib[:] = (np.diag([0.011253, 0.036203, 0.042673]) + 1e-3 * (A3 @ A3.T)).ravel().tolist()
    p.kff[:] = [0.3, -0.2, 0.15, 0.5, -0.4, 0.7]
    p.kp_p[:] = [80.0, 120.0, 150.0]
    p.fzmin, p.fzmax = 0.0, 90.0
    p.clamp_tau, p.tau_min, p.tau_max = 1, -6.0, 5.0
    _check(host, p, states.generate_states(3000, 17, profile="stress", masks="mixed"), 1e-7)


def test_core_refuses_general_W(host, params06):
    p = params06.copy()
    W = 1e-5 * np.eye(12)
    W[0, 1] = W[1, 0] = 1e-7
    p.W[:] = W.ravel().tolist()
    rc, _ = host(p, states.generate_states(4, 1))
    assert rc == -1  # qpb_create routes such parameter sets to the half-warp kernel


def test_core_iteration_limit_and_bad_input(host, params06):
    S = states.generate_states(256, 3, profile="stress")
    S["xdot"][5, 1] = np.nan
    S["Rwb"][9, 4] = np.inf
    rc, out = host(params06, S)
    assert list(np.nonzero(out["status"])[0]) == [5, 9] and not out["grf_body"][[5, 9]].any()
    p = params06.copy()
    p.max_iter = 3
    rc, out3 = host(p, S)
    over = out3["status"] == 1
    assert over.any() and (out3["iters"][over] == 3).all() and not out3["grf_body"][over].any()


def test_core_warm_start_is_the_reference_hotstart(host, params06):
    """Feeding a QP its own final working set back (the reference's SQProblem::hotstart between ticks,
    balance_controller.cpp:177-202) must give the same answer in zero working-set changes; a perturbed state started
    from the previous working set must agree with its cold solve; a nonsense hint must fall back to a cold start."""
    S = states.generate_states(4000, 20260103, masks="mixed")
    rc, cold = host(params06, S)
    words = cold["pad"][:, :4].copy().view("<u4")[:, 0]
    assert (words >> 31).all()
    W = S.copy()
    W["pad"][:, :4] = cold["pad"][:, :4]
    rc, warm = host(params06, W)
    assert (warm["iters"] == 0).all()
    assert rel_err(warm["grf_body"], cold["grf_body"]) <= 1e-7
    # next tick: the state moved a little
    rng = np.random.default_rng(3)
    N = S.copy()
    N["x"] += rng.normal(0, 2e-4, N["x"].shape)
    N["xdot"] += rng.normal(0, 2e-3, N["xdot"].shape)
    N["w"] += rng.normal(0, 2e-3, N["w"].shape)
    rc, ncold = host(params06, N)
    N["pad"][:, :4] = cold["pad"][:, :4]
    rc, nwarm = host(params06, N)
    assert np.array_equal(nwarm["status"], ncold["status"])
    assert rel_err(nwarm["grf_body"], ncold["grf_body"]) <= 1e-7
    assert nwarm["iters"].mean() <= 2.0 < ncold["iters"].mean()
    # garbage hints
    G = S.copy()
    G["pad"][:, :4] = np.frombuffer(rng.integers(0, 2**32, len(S), dtype=np.uint32).tobytes(), dtype=np.uint8).reshape(-1, 4)
    G["pad"][:, 3] |= 0x80
    rc, garb = host(params06, G)
    assert np.array_equal(garb["status"], cold["status"])
    assert rel_err(garb["grf_body"], cold["grf_body"]) <= 1e-7
