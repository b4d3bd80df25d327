"""10-step convex-MPC QP (BASELINE config 4).

PARITY UNPINNED: the reference has no code for this path (README.md:22-26), so there is nothing of the reference's
to compare against.  What is checked: the oracle (oracle/mpc_oracle.c, literal column-by-column simulation of the
prediction matrices + dense Goldfarb-Idnani in the reference's two-sided row form) passes the solver-free KKT
certificate; an independent numpy restatement of the CUDA kernel's closed-form assembly agrees with it; and on the
GPU the CUDA path agrees with the oracle to 1e-5 relative (the tolerance north_star states for forces)."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import rel_err
from quadruped_control_b200 import lib
from quadruped_control_b200.records import MPC_OUT_DTYPE, MPC_REC_DTYPE, MpcParams, default_mpc_params
from quadruped_control_b200.states import generate_mpc

TOL = 1e-5  # relative, floor 1 N (SURVEY.md 8d metric)
NCPU = os.cpu_count() or 1


@pytest.fixture(scope="module")
def mpc_params():
    return default_mpc_params()


# ---------------------------------------------------------------------------------------------------- CPU --
def test_mpc_params_mirror_and_layouts(built, tmp_path):
    import subprocess

    import oracle

    assert bytes(lib.default_mpc_params()) == bytes(default_mpc_params()) == bytes(oracle.mpc_default_params())
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "qpb200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(qpb_mpc_params), sizeof(qpb_mpc_rec),'
        " sizeof(qpb_mpc_out_rec), offsetof(qpb_mpc_rec, xref), offsetof(qpb_mpc_rec, r), offsetof(qpb_mpc_rec, contact),"
        " offsetof(qpb_mpc_out_rec, status));return 0;}\n"
    )
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got[0] == ctypes.sizeof(MpcParams)
    assert got[1] == MPC_REC_DTYPE.itemsize == 2176 and got[2] == MPC_OUT_DTYPE.itemsize == 1024
    assert got[3] == MPC_REC_DTYPE.fields["xref"][1] == 104 and got[4] == MPC_REC_DTYPE.fields["r"][1] == 1144
    assert got[5] == MPC_REC_DTYPE.fields["contact"][1] == 2104 and got[6] == MPC_OUT_DTYPE.fields["status"][1] == 960


def test_mpc_generator_is_index_addressable():
    a = generate_mpc(64, 20260104)
    b = generate_mpc(16, 20260104, lo=40)
    assert a[40:56].tobytes() == b.tobytes()
    gaits = a["contact"].sum(axis=(1, 2))
    assert set(np.unique(gaits)) <= {20, 30, 40} and len(np.unique(gaits)) == 3
    assert (a["x0"][:, 12] == -9.81).all() and np.isfinite(a["r"]).all()


def test_mpc_oracle_passes_kkt_certificate(built, mpc_params):
    import oracle
    from oracle import kkt

    R = np.concatenate([generate_mpc(24, 20260104), generate_mpc(8, 7, scale=3.0)])
    out = oracle.mpc_batch(mpc_params, R, NCPU)
    assert (out["status"] == 0).all()
    for i in range(len(R)):
        q = oracle.mpc_assemble(mpc_params, R[i])
        cert = kkt.certificate(q["Q"], q["c"], q["C"], q["lb"], q["ub"], out["U"][i])
        assert cert["infeas"] <= 1e-9 and cert["stat_rel"] <= 1e-10, (i, cert["infeas"], cert["stat_rel"])
        swing = np.repeat(R[i]["contact"].reshape(40) == 0, 3)
        assert np.abs(out["U"][i][swing]).max(initial=0.0) <= 1e-9


def test_mpc_oracle_is_equivariant_under_the_left_right_mirror(built, mpc_params):
    """The model, the weights and the pyramids are symmetric under y -> -y (roll, yaw, omega_x, omega_z and every y
    component change sign, legs RL <-> RR and FL <-> FR swap): the mirrored problem must return the mirrored forces.
    A sign or index slip anywhere in the prediction matrices, the cross products or the leg order breaks this."""
    import oracle

    R = np.concatenate([generate_mpc(32, 20260104), generate_mpc(16, 11, scale=2.5)])
    ref = oracle.mpc_batch(mpc_params, R, NCPU)
    assert (ref["status"] == 0).all()
    sgn = np.array([-1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, 1], dtype=np.float64)  # [roll pitch yaw | p | omega | v | g]
    perm = [2, 3, 0, 1]
    my = np.array([1.0, -1.0, 1.0])
    Rm = R.copy()
    Rm["x0"] = R["x0"] * sgn
    Rm["xref"] = R["xref"] * sgn
    Rm["r"] = R["r"][:, :, perm, :] * my
    Rm["contact"] = R["contact"][:, :, perm]
    out = oracle.mpc_batch(mpc_params, Rm, NCPU)
    assert (out["status"] == 0).all()
    want = (ref["U"].reshape(-1, 10, 4, 3)[:, :, perm, :] * my).reshape(-1, 120)
    assert rel_err(out["U"], want) <= 1e-7


def test_mpc_oracle_agrees_with_closed_form_restatement(built, mpc_params):
    """The closed-form condensed Hessian the CUDA kernel uses (numpy restatement in tests/tools/prototypes/proto_mpc.py)
    against the oracle's literal simulation of the prediction matrices, and the operator-form active-set loop against
    the oracle's QR-form solver."""
    import oracle

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "tools", "prototypes"))
    import proto_mpc

    R = generate_mpc(12, 20260104)
    ref = oracle.mpc_batch(mpc_params, R, NCPU)
    for i in range(len(R)):
        q = oracle.mpc_assemble(mpc_params, R[i])
        H, g, vmap, ns = proto_mpc.closed_form(mpc_params, R[i])
        assert np.abs(H - q["Q"][np.ix_(vmap, vmap)]).max() <= 1e-12 * np.abs(q["Q"]).max()
        assert np.abs(g - q["c"][vmap]).max() <= 1e-12 * (1.0 + np.abs(q["c"]).max())
        st, f, _ = proto_mpc.solve(mpc_params, H, g, ns)
        U = np.zeros(120)
        U[vmap] = f
        assert st == 0 and rel_err(U[None], ref["U"][i][None]) <= 1e-9


def test_mpc_oracle_reproduces_golden_fixture(built, mpc_params):
    import oracle

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mpc_golden.npz"))
    R = np.ascontiguousarray(g["recs"]).view(MPC_REC_DTYPE).reshape(-1)
    out = oracle.mpc_batch(mpc_params, R, NCPU)
    assert np.array_equal(out["status"], g["status"]) and rel_err(out["U"], g["U"]) <= 1e-10


def test_mpc_oracle_rejects_nonfinite(built, mpc_params):
    import oracle

    R = generate_mpc(2, 3)
    R["xref"][1, 4, 2] = np.nan
    out = oracle.mpc_batch(mpc_params, R, 1)
    assert out["status"][0] == 0 and out["status"][1] == 2 and not out["U"][1].any()


def test_mpc_create_validates_parameters(built):
    L = lib.load()

    def rc_of(mut):
        p = default_mpc_params()
        mut(p)
        h = ctypes.c_void_p()
        rc = L.qpb_mpc_create(ctypes.byref(p), 0, ctypes.byref(h))
        if rc == 0:
            L.qpb_mpc_destroy(h)
        return rc

    for mut in (lambda p: setattr(p, "mu", 0.0), lambda p: setattr(p, "fzmin", 200.0), lambda p: setattr(p, "alpha", 0.0),
                lambda p: setattr(p, "dt", -1.0), lambda p: setattr(p, "max_iter", 0), lambda p: setattr(p, "mass", float("nan")),
                lambda p: p.Lw.__setitem__(3, -1.0), lambda p: p.Ib.__setitem__(0, 0.0)):
        assert rc_of(mut) == -2  # QPB_ERR_BAD_PARAMS, decided before any CUDA call


# ---------------------------------------------------------------------------------------------------- GPU --
@pytest.fixture(scope="module")
def mpc_solver(built, mpc_params):
    s = lib.MpcSolver(mpc_params, device=0)
    yield s
    s.close()


def _feasible(p, R, out, tol=1e-6):
    U = out["U"].reshape(-1, 10, 4, 3)
    st = R["contact"].astype(bool)
    fx, fy, fz = U[..., 0], U[..., 1], U[..., 2]
    ok = np.abs(U[~st]).max(initial=0.0) == 0.0
    ok &= (np.abs(fx[st]) <= p.mu * fz[st] + tol).all() and (np.abs(fy[st]) <= p.mu * fz[st] + tol).all()
    ok &= (fz[st] >= p.fzmin - tol).all() and (fz[st] <= p.fzmax + tol).all()
    return bool(ok)


@pytest.mark.gpu
@pytest.mark.parametrize("gaits,scale,seed", [("mixed", 1.0, 20260104), ("stand", 1.0, 11), ("trot", 1.0, 12), ("crawl", 1.0, 13),
                                              ("mixed", 3.0, 14), ("stand", 6.0, 15)])
def test_mpc_gpu_matches_oracle(mpc_solver, mpc_params, gaits, scale, seed):
    import oracle

    R = generate_mpc(384, seed, gaits=gaits, scale=scale)
    out = mpc_solver.solve_host(R)
    ref = oracle.mpc_batch(mpc_params, R, NCPU)
    assert (ref["status"] == 0).all()
    assert (out["status"] == ref["status"]).all(), np.bincount(out["status"])
    err = rel_err(out["U"], ref["U"])
    assert err <= TOL, err
    assert _feasible(mpc_params, R, out)


@pytest.mark.gpu
def test_mpc_gpu_ragged_and_disturbed_batch_matches_oracle(built, mpc_params):
    """Every gait, two disturbance scales and ragged / empty contact patterns in one batch."""
    import oracle

    s = lib.MpcSolver(mpc_params, device=0)
    R = np.concatenate([generate_mpc(300, 71), generate_mpc(150, 72, scale=4.0), generate_mpc(60, 73, gaits="trot")])
    R["contact"][5] = 0
    R["contact"][6, ::3, 1] = 0
    out = s.solve_host(R)
    s.close()
    ref = oracle.mpc_batch(mpc_params, R, NCPU)
    assert (out["status"] == ref["status"]).all() and (ref["status"] == 0).all()
    assert rel_err(out["U"], ref["U"]) <= TOL
    assert _feasible(mpc_params, R, out)


@pytest.mark.gpu
def test_mpc_gpu_random_parameter_sets(built):
    """Fuzz over controller parameters: friction, force limits, mass, a full inertia matrix, horizon step, state weights
    (some zero) and force weights down to 1e-7 (condition numbers up to ~1e8)."""
    import oracle

    rng = np.random.default_rng(2026)
    worst = 0.0
    for trial in range(8):
        p = default_mpc_params()
        p.mu = float(rng.uniform(0.2, 1.2))
        p.fzmin = float(rng.choice([0.0, 5.0, 10.0]))
        p.fzmax = float(rng.uniform(60.0, 300.0))
        p.mass = float(rng.uniform(5.0, 40.0))
        A = rng.normal(size=(3, 3)) * 0.01
        Ib = np.diag([0.011253, 0.036203, 0.042673]) * rng.uniform(0.5, 5.0) + A @ A.T
        p.Ib[:] = Ib.ravel().tolist()
        p.dt = float(rng.uniform(0.01, 0.06))
        Lw = np.array(default_mpc_params().Lw[:]) * rng.uniform(0.1, 10.0, size=13)
        Lw[rng.integers(0, 12)] = 0.0
        p.Lw[:] = Lw.tolist()
        p.alpha = float(10.0 ** rng.uniform(-7, -3))
        R = generate_mpc(96, 500 + trial, params=p, scale=float(rng.choice([1.0, 2.5])))
        s = lib.MpcSolver(p, device=0)
        out = s.solve_host(R)
        s.close()
        ref = oracle.mpc_batch(p, R, NCPU)
        assert (ref["status"] == 0).all() and (out["status"] == 0).all(), (trial, np.bincount(out["status"]))
        err = rel_err(out["U"], ref["U"])
        worst = max(worst, err)
        assert err <= TOL, (trial, err, p.alpha)
        assert _feasible(p, R, out)
    print("worst rel err over parameter sets", worst)


@pytest.mark.gpu
def test_mpc_gpu_passes_kkt_certificate(mpc_solver, mpc_params):
    import oracle
    from oracle import kkt

    R = generate_mpc(24, 99, scale=2.0)
    out = mpc_solver.solve_host(R)
    for i in range(len(R)):
        q = oracle.mpc_assemble(mpc_params, R[i])
        cert = kkt.certificate(q["Q"], q["c"], q["C"], q["lb"], q["ub"], out["U"][i])
        assert out["status"][i] == 0 and cert["infeas"] <= 1e-8 and cert["stat_rel"] <= 1e-9, (i, cert["infeas"], cert["stat_rel"])


@pytest.mark.gpu
def test_mpc_gpu_matches_golden_fixture(mpc_solver):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mpc_golden.npz"))
    R = np.ascontiguousarray(g["recs"]).view(MPC_REC_DTYPE).reshape(-1)
    out = mpc_solver.solve_host(R)
    assert np.array_equal(out["status"], g["status"]) and rel_err(out["U"], g["U"]) <= TOL


@pytest.mark.gpu
def test_mpc_gpu_edge_cases(mpc_solver, mpc_params):
    import oracle

    assert len(mpc_solver.solve_host(np.zeros(0, dtype=MPC_REC_DTYPE))) == 0
    R = generate_mpc(8, 5)
    R["contact"][0] = 0                     # flight phase over the whole horizon: no variables at all
    R["contact"][1] = 0
    R["contact"][1, 3, 2] = 1               # a single stance foot-step
    R["x0"][2, 7] = np.inf                  # non-finite state
    R["r"][3, 9, 3, 2] = np.nan
    R["contact"][4, :, :] = 1
    R["contact"][4, ::2, 0] = 0             # ragged pattern: one leg toggling every step
    out = mpc_solver.solve_host(R)
    ref = oracle.mpc_batch(mpc_params, R, NCPU)
    assert list(out["status"]) == list(ref["status"]) == [0, 0, 2, 2, 0, 0, 0, 0]
    assert not out["U"][0].any() and not out["U"][2].any() and not out["U"][3].any()
    assert rel_err(out["U"], ref["U"]) <= TOL


@pytest.mark.gpu
def test_mpc_gpu_iteration_limit(built, mpc_params):
    p = mpc_params.copy()
    p.max_iter = 2
    s = lib.MpcSolver(p, device=0)
    R = generate_mpc(64, 20260104, scale=3.0)
    out = s.solve_host(R)
    s.close()
    assert (out["status"] == 1).any() and set(np.unique(out["status"])) <= {0, 1}
    assert not out["U"][out["status"] == 1].any() and (out["iters"] <= 2).all()


@pytest.mark.gpu
def test_mpc_gpu_packed_equals_host_and_is_deterministic(mpc_solver):
    import torch

    R = generate_mpc(1000, 42)
    a = mpc_solver.solve_host(R)
    d_in = torch.from_numpy(R.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(len(R) * MPC_OUT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        mpc_solver.solve_packed(d_in, d_out, len(R), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        b = d_out.cpu().numpy().view(MPC_OUT_DTYPE)
        assert a.tobytes() == b.tobytes()


@pytest.mark.gpu
def test_mpc_gpu_full_size_properties(mpc_solver, mpc_params):
    """BASELINE config 4 at full size (65 536 QPs): every answer feasible, statuses OK, a spot sample against the oracle
    and a chunk-invariance check (the same records solved in a different batch give the same bytes)."""
    import oracle

    R = generate_mpc(65536, 20260104)
    out = mpc_solver.solve_host(R)
    assert (out["status"] == 0).all()
    assert _feasible(mpc_params, R, out)
    idx = np.random.default_rng(0).choice(len(R), 256, replace=False)
    ref = oracle.mpc_batch(mpc_params, R[idx], NCPU)
    assert rel_err(out["U"][idx], ref["U"]) <= TOL
    again = mpc_solver.solve_host(R[idx])
    assert again.tobytes() == out[idx].tobytes()
