"""Solver-free KKT certificate for the balance QP.  TEST INFRASTRUCTURE ONLY.

For  min 1/2 x'Qx + c'x  s.t.  lb <= Cx <= ub  (Q > 0) a point x is THE minimiser iff it is
feasible and Qx + c = C' lam with lam_i >= 0 where row i sits at its lower bound, <= 0 at its
upper bound, free on equality rows and 0 elsewhere.  The certificate finds the best such lam by
non-negative least squares over the rows that are active at x and reports
  * primal infeasibility (absolute),
  * the stationarity residual r = Qx + c - C' lam (infinity norm),
  * a rigorous distance bound  ||x - x*||_2 <= (||r||_2 + sqrt(2 lmin * slack_gap)) / lmin  is
    not attempted; instead ``dist_bound`` = ||r||_2 / lmin(Q) which holds when x is feasible and
    lam is complementary on rows that are exactly active (strong convexity).
Because Q = 2(A'SA + W) has lmin >= 2 lmin(W) = 2e-5, a residual of 1e-10 pins x to 5e-6 N.
"""
import numpy as np
from scipy.optimize import nnls


def certificate(Q, c, C, lb, ub, x, act_tol=1e-7):
    Q, c, C, lb, ub, x = (np.asarray(v, dtype=np.float64) for v in (Q, c, C, lb, ub, x))
    Cx = C @ x
    scale = 1.0 + np.abs(x).max()
    infeas = max(0.0, float((lb - Cx).max()), float((Cx - ub).max()))
    g = Q @ x + c
    cols, meta = [], []
    for i in range(C.shape[0]):
        at_lb = abs(Cx[i] - lb[i]) <= act_tol * scale
        at_ub = abs(Cx[i] - ub[i]) <= act_tol * scale
        if at_lb or lb[i] == ub[i]:
            cols.append(C[i])
            meta.append((i, +1.0))
        if at_ub or lb[i] == ub[i]:
            cols.append(-C[i])
            meta.append((i, -1.0))
    lam = np.zeros(C.shape[0])
    if cols:
        N = np.array(cols).T
        # column scaling keeps NNLS well conditioned
        y, _ = nnls(N, g, maxiter=50 * N.shape[1] + 50)
        r = g - N @ y
        for (i, sgn), yi in zip(meta, y):
            lam[i] += sgn * yi
    else:
        r = g
    lmin = float(np.linalg.eigvalsh(Q)[0])
    return dict(
        infeas=infeas,
        stat_inf=float(np.abs(r).max()),
        stat_rel=float(np.abs(r).max() / (1.0 + np.abs(g).max() + np.abs(c).max())),
        dist_bound=float(np.linalg.norm(r) / lmin),
        lam=lam,
        n_active=int(np.count_nonzero(np.abs(lam) > 0)),
    )
