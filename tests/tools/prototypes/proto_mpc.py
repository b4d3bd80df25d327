"""Numpy prototype of the CUDA MPC kernel's arithmetic (closed-form condensed Hessian on the compacted stance
variables + whitened operator-form Goldfarb-Idnani), checked against the oracle.  Developer tool, not a test."""
import sys

import numpy as np

sys.path.insert(0, ".")
import oracle  # noqa: E402
from quadruped_control_b200.records import default_mpc_params  # noqa: E402
from quadruped_control_b200.states import generate_mpc  # noqa: E402


def closed_form(p, rec):
    dt, m, al = p.dt, p.mass, p.alpha
    Lw = np.array(p.Lw[:])
    Ibinv = np.linalg.inv(np.array(p.Ib[:]).reshape(3, 3))
    contact = rec["contact"].reshape(40)
    sf = [c for c in range(40) if contact[c]]
    ns = len(sf)
    n = 3 * ns
    psi = rec["xref"][:, 2]
    cs, sn = np.cos(psi), np.sin(psi)
    Cs, Ss = np.cumsum(cs), np.cumsum(sn)
    G = np.zeros((n, 3))
    step = np.zeros(n, dtype=int)
    comp = np.zeros(n, dtype=int)
    for a in range(n):
        c = sf[a // 3]
        k, foot, cc = c // 4, c % 4, a % 3
        Rz = np.array([[cs[k], -sn[k], 0], [sn[k], cs[k], 0], [0, 0, 1]])
        Iinv = Rz @ Ibinv @ Rz.T
        e = np.zeros(3)
        e[cc] = 1
        G[a] = Iinv @ np.cross(rec["r"][k, foot], e)
        step[a], comp[a] = k, cc
    # Theta table
    Th = np.zeros((n, 10, 3))
    for a in range(n):
        j = step[a]
        for k in range(j + 1, 10):
            C, S = Cs[k] - Cs[j], Ss[k] - Ss[j]
            g = G[a]
            Th[a, k] = dt * dt * np.array([C * g[0] + S * g[1], -S * g[0] + C * g[1], (k - j) * g[2]])
    # free response error
    x0 = rec["x0"]
    E = np.zeros((10, 12))
    for k in range(10):
        w0 = x0[6:9]
        th = x0[0:3] + dt * np.array([Cs[k] * w0[0] + Ss[k] * w0[1], -Ss[k] * w0[0] + Cs[k] * w0[1], (k + 1) * w0[2]])
        v = x0[9:12] + np.array([0, 0, (k + 1) * dt * x0[12]])
        pp = x0[3:6] + (k + 1) * dt * x0[9:12] + np.array([0, 0, dt * dt * x0[12] * k * (k + 1) / 2])
        free = np.concatenate([th, pp, w0, v])
        E[k] = Lw[:12] * (free - rec["xref"][k, :12])
    H = np.zeros((n, n))
    g = np.zeros(n)
    for a in range(n):
        ja, ca = step[a], comp[a]
        for b in range(a + 1):
            jb, cb = step[b], comp[b]
            cnt = 10 - ja
            t = sum((Lw[0:3] * Th[a, k]) @ Th[b, k] for k in range(ja + 1, 10))
            t += dt * dt * ((Lw[6:9] * G[a]) @ G[b]) * cnt
            if ca == cb:
                t += (dt / m) ** 2 * Lw[9 + ca] * cnt
                t += (dt * dt / m) ** 2 * Lw[3 + ca] * sum((k - ja) * (k - jb) for k in range(ja, 10))
            if a == b:
                t += al
            H[a, b] = H[b, a] = 2 * t
        t = 0.0
        for k in range(ja, 10):
            t += Th[a, k] @ E[k, 0:3] + (k - ja) * dt * dt / m * E[k, 3 + ca] + dt * (G[a] @ E[k, 6:9]) + dt / m * E[k, 9 + ca]
        g[a] = 2 * t
    vmap = np.array([12 * (sf[a // 3] // 4) + 3 * (sf[a // 3] % 4) + a % 3 for a in range(n)])
    return H, g, vmap, ns


def rows(p, ns):
    """6 one-sided rows per stance foot-step: n.f >= b, (var offset, coefficient) pairs."""
    mu = p.mu
    out = []
    for c in range(ns):
        v = 3 * c
        out += [([(v, -1.0), (v + 2, mu)], 0.0), ([(v + 1, -1.0), (v + 2, mu)], 0.0), ([(v + 1, 1.0), (v + 2, mu)], 0.0),
                ([(v, 1.0), (v + 2, mu)], 0.0), ([(v + 2, 1.0)], p.fzmin), ([(v + 2, -1.0)], -p.fzmax)]
    return out


def solve(p, H, g, ns, max_iter=1000):
    n = 3 * ns
    L = np.linalg.cholesky(H)
    X = np.linalg.inv(L)  # Linv
    f = X.T @ (-(X @ g))
    R = rows(p, ns)
    m = len(R)
    Nd = np.zeros((m, n))
    bd = np.zeros(m)
    for j, (terms, b) in enumerate(R):
        for v, cf in terms:
            Nd[j, v] = cf
        bd[j] = b
    act = np.zeros(m, dtype=int)
    A, u = [], []
    Ns = np.zeros((0, n))
    it = 0
    while True:
        s = Nd @ f - bd
        viol = (act == 0) & (s < -1e-9 * (1 + np.abs(bd)))
        if not viol.any():
            return 0, f, it
        pidx = int(np.argmin(np.where(viol, s, np.inf)))
        up = 0.0
        sp = s[pidx]
        while True:
            it += 1
            if it > max_iter:
                return 1, f, it
            nt = X @ Nd[pidx]
            r = Ns @ nt
            w = Nd[A].T @ r if A else np.zeros(n)
            zt = nt - X @ w
            zeta = nt @ zt
            nn = nt @ nt
            dep = not (zeta > 1e-13 * nn)
            t2 = np.inf if dep else -sp / zeta
            t1, ks = np.inf, -1
            for k in range(len(A)):
                if r[k] > 0:
                    tt = max(u[k], 0.0) / r[k]
                    if tt < t1:
                        t1, ks = tt, k
            if dep and ks < 0:
                act[pidx] = 2
                break
            full = (not dep) and t2 <= t1
            t = t2 if full else t1
            if not dep:
                f = f + t * (X.T @ zt)
                sp += t * zeta
            u = [uk - t * rk for uk, rk in zip(u, r)]
            up += t
            if full:
                Ns = np.vstack([Ns - np.outer(r, zt) / zeta, zt / zeta])
                A.append(pidx)
                u.append(up)
                act[pidx] = 1
                break
            nu = Ns[ks].copy()
            d = Ns @ nu
            Ns = Ns - np.outer(d / d[ks], nu)
            act[A[ks]] = 0
            last = len(A) - 1
            Ns[ks] = Ns[last]
            A[ks] = A[last]
            u[ks] = u[last]
            Ns = Ns[:last]
            A.pop()
            u.pop()


if __name__ == "__main__":
    p = default_mpc_params()
    R = generate_mpc(48, 20260104)
    ref = oracle.mpc_batch(p, R, 8)
    worstH = worstU = 0.0
    for i in range(len(R)):
        q = oracle.mpc_assemble(p, R[i])
        H, g, vmap, ns = closed_form(p, R[i])
        Hd = q["Q"][np.ix_(vmap, vmap)]
        worstH = max(worstH, np.abs(H - Hd).max() / np.abs(Hd).max(), np.abs(g - q["c"][vmap]).max() / (1 + np.abs(q["c"]).max()))
        st, f, it = solve(p, H, g, ns)
        U = np.zeros(120)
        U[vmap] = f
        err = np.abs(U - ref["U"][i]).max() / max(np.abs(ref["U"][i]).max(), 1.0)
        worstU = max(worstU, err)
        print(i, ns, st, it, ref["iters"][i], f"{err:.2e}")
    print("worst H/g rel", worstH, "worst U rel", worstU)
