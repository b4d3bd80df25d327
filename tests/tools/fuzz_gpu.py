"""Randomised parameter fuzz: CUDA path (default kernel choice and both warp mappings) vs the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, oracle
from quadruped_control_b200 import default_params, states, lib
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
nper = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 2026)
worst = 0.0; worst_desc = ''; fails = 0; ncpu = os.cpu_count() or 1
for t in range(trials):
    p = default_params(float(np.exp(rng.uniform(np.log(0.02), np.log(3.0)))))
    p.mass = float(rng.uniform(2.0, 80.0))
    fz = np.sort(rng.uniform(0.0, 400.0, 2)); 
    if rng.random() < 0.2: fz[0] = 0.0
    if rng.random() < 0.1: fz[1] = fz[0]
    p.fzmin, p.fzmax = float(fz[0]), float(max(fz[1], 1.0))
    if p.fzmin > p.fzmax: p.fzmin = p.fzmax
    A6 = rng.normal(size=(6, 6)); A12 = rng.normal(size=(12, 12)); A3 = rng.normal(size=(3, 3))
    sS, sW = rng.choice([0.0, 0.05, 0.5]), rng.choice([0.0, 0.2, 2.0])
    S6 = np.diag(np.exp(rng.uniform(np.log(0.2), np.log(50.0), 6))) + sS * (A6 @ A6.T)
    w = float(np.exp(rng.uniform(np.log(1e-6), np.log(1e-2))))
    W12 = w * (np.eye(12) + sW * (A12 @ A12.T) / 12.0)
    Ib = np.diag(rng.uniform(0.005, 0.5, 3)) + 1e-3 * rng.choice([0.0, 1.0]) * (A3 @ A3.T)
    p.S[:] = S6.ravel().tolist(); p.W[:] = W12.ravel().tolist(); p.Ib[:] = Ib.ravel().tolist()
    p.kff[:] = rng.uniform(-0.5, 0.5, 6).tolist()
    p.kp_p[:] = rng.uniform(10, 300, 3).tolist(); p.kd_p[:] = rng.uniform(1, 100, 3).tolist()
    p.kp_w[:] = rng.uniform(100, 8000, 3).tolist(); p.kd_w[:] = rng.uniform(10, 800, 3).tolist()
    prof = str(rng.choice(["default", "light", "stress"])); masks = str(rng.choice(["all4", "mixed"]))
    Sin = states.generate_states(nper, int(rng.integers(1, 1 << 30)), profile=prof, masks=masks, params=p)
    if rng.random() < 0.3:
        Sin["x_d"][:, 2] -= rng.uniform(0, 0.5, nper)
    ref = oracle.control_batch(p, Sin, ncpu)
    # auto: the dispatch as shipped; three / one: the three-pass and the one-launch range-space paths forced at this size
    # (they apply when W = w I and fzmin >= 0, else the half-warp kernel runs); warm: the same records carrying the
    # working sets the previous mode returned; 2 / 1: half-warp and one-warp kernels
    prev = None
    for mode in ("auto", "three", "one", "warm", "2", "1"):
        for k in ("QPB_QPS_PER_WARP", "QPB_TPQ_MIN_N", "QPB_TPQ_ONE_MAX"):
            os.environ.pop(k, None)
        if mode in ("2", "1"):
            os.environ["QPB_QPS_PER_WARP"] = mode
        elif mode == "three":
            os.environ["QPB_TPQ_MIN_N"] = "0"
        elif mode == "one":
            os.environ["QPB_TPQ_MIN_N"] = str(1 << 31); os.environ["QPB_TPQ_ONE_MAX"] = str(1 << 30)
        X = Sin
        if mode == "warm":
            X = Sin.copy(); X["pad"][:, :4] = prev["pad"][:, :4]
        sol = lib.BalanceSolver(p); out = sol.control_host(X); sol.close()
        prev = out
        mism = int((out["status"] != ref["status"]).sum())
        ok = (out["status"] == 0) & (ref["status"] == 0)
        err = float((np.abs(out["grf_body"][ok] - ref["grf_body"][ok]).max(axis=1) / np.maximum(np.abs(ref["grf_body"][ok]).max(axis=1), 1)).max()) if ok.any() else 0.0
        errt = float((np.abs(out["tau"][ok] - ref["tau"][ok]).max(axis=1) / np.maximum(np.abs(ref["tau"][ok]).max(axis=1), 1)).max()) if ok.any() else 0.0
        if max(err, errt) > worst:
            worst = max(err, errt)
            worst_desc = f"trial {t} mode {mode}: mu={p.mu:.3g} mass={p.mass:.3g} fz=[{p.fzmin:.3g},{p.fzmax:.3g}] w={w:.2e} sS={sS} sW={sW} {prof}/{masks} err {err:.2e} {errt:.2e}"
        if mism or err > 1e-5 or errt > 1e-5:
            fails += 1
            print(f"FAIL trial {t} mode {mode}: mu={p.mu:.3g} mass={p.mass:.3g} fz=[{p.fzmin:.3g},{p.fzmax:.3g}] w={w:.2e} sS={sS} sW={sW} {prof}/{masks} "
                  f"status mism {mism} (gpu {np.bincount(out['status'], minlength=3)}, ref {np.bincount(ref['status'], minlength=3)}) err {err:.2e} {errt:.2e} iters max {out['iters'].max()}")
print("worst case:", worst_desc)
print(f"fuzz: {trials} parameter sets x {nper} states x 6 kernel choices, failures {fails}, worst rel err {worst:.2e}")
