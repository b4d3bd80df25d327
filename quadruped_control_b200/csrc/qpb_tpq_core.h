// qpb_tpq_core.h -- the balance QP solved by ONE THREAD, in the 6-dimensional range space of the wrench map.
//
// Valid when the force regulariser is a multiple of the identity, W = w I -- the reference's only configuration
// (commander_node.cpp:289, 305; Hessian at balance_controller.cpp:152-153).  General W takes the half-warp kernel.
//
//   min (A f - b)' S (A f - b) + w f'f     s.t. per stance leg  |fx| <= mu fz, |fy| <= mu fz, fzmin <= fz <= fzmax
//
// with A = [A_0 .. A_3], A_i = [I; [r_i]x] (balance_controller.cpp:237-272).  Every inequality row touches one leg,
// so a working set is a choice of face per leg, and on a fixed set of faces the minimiser is
//   f_i = p_i - (1/w) Pi_i A_i' y,     (w S^-1 + sum_i A_i Pi_i A_i') y = w (A p - b)
// (Pi_i: 3x3 projector onto the tangent space of leg i's face, p_i: the point of the face closest to 0) -- a 6x6 SPD
// system whatever the working set.  The solver is Goldfarb-Idnani's dual active-set method (the same method, selection
// rule and tolerances as the half-warp kernel, so the iteration path is the same), but each step direction
//   z = (1/2w) (Pi n_p - 2 Pi A' yh),   G yh = (1/2) A_l Pi_l n_p,      N r = n_p - Q z  (leg by leg, closed form)
// is solved afresh from a 6x6 Cholesky factor: no operator is updated from step to step, so no error accumulates
// beyond the iterate itself.  State per QP: f (12), multipliers (12), lever arms (12), one sign per row group.
//
// Working sets are restricted to at most one row per group (x, y, z) and leg -- the faces of the truncated pyramid.
// Goldfarb-Idnani may add any violated row, so the selection rule simply never picks the second row of a group that is
// already active on a leg: that row (fx + mu fz >= 0 with -fx + mu fz = 0 active, say) can only be violated while
// fz < 0, and then the leg's fz >= fzmin row is violated too and goes first.  This needs fzmin >= 0; qpb_create sends
// parameter sets with fzmin < 0 (a foot pulling on the ground) to the general half-warp kernel.  When the loop ends,
// polish() recomputes the minimiser on the final faces from scratch and re-checks every row.
//
// Warm start (the reference's hotstart, balance_controller.cpp:177-202): a previous working set may be supplied as a
// 24-bit word; the solver computes the minimiser on those faces and, if all its multipliers are non-negative, continues
// from there (that is a valid dual-feasible starting pair); otherwise it cold-starts.
//
// Everything here compiles for the host as well: tests/ build it with g++ to check the algorithm against the oracle
// on the CPU.  The product path is the CUDA kernel in qpb_tpq.cuh only.
#pragma once

#include "qpb_stages.h"

namespace qpb {
namespace tpq {


// Derived once at qpb_create; travels as a kernel argument (constant bank).
struct FastParams {
  double w, inv_w;   // W = w I
  double wSinv[21];  // w S^-1, packed lower triangle: (i, j <= i) at i (i + 1) / 2 + j
  double mu, k1, k2; // k1 = 1 / (1 + mu^2), k2 = 1 / (1 + 2 mu^2)
  double fzmin, fzmax;
  double ntol_z;     // violation tolerance of the fz rows: -1e-9 (1 + max(|fzmin|, |fzmax|))
  int max_iter;      // working-set changes allowed (params.max_iter)
};

#define ix(i, j) ((i) * ((i) + 1) / 2 + (j))  // packed lower triangle

// Face of one leg: Pi = diag(ax, ay, 0) + kap d d', d = (dx, dy, 1).
struct Face {
  double ax, ay, dx, dy, kap;
};

// sx, sy, sz: coefficient of the active row of the group on its own axis (0 = none):
//   x rows (sx, 0, mu), y rows (0, sy, mu), z rows (0, 0, sz)   [A rows: sx = sy = -1, sz = +1; B rows the opposite]
QPB_HD Face face_of(double sx, double sy, double sz, bool stance, const FastParams& K) {
  const bool X = sx != 0.0, Y = sy != 0.0, Z = sz != 0.0;
  Face F;
  F.ax = (X || !stance) ? 0.0 : 1.0;
  F.ay = (Y || !stance) ? 0.0 : 1.0;
  F.dx = -sx * K.mu;
  F.dy = -sy * K.mu;
  F.kap = (Z || !stance) ? 0.0 : ((X && Y) ? K.k2 : ((X || Y) ? K.k1 : 1.0));
  return F;
}

QPB_HD void face_proj(const Face& F, const double (&g)[3], double (&o)[3]) {
  const double dg = F.kap * fma(F.dx, g[0], fma(F.dy, g[1], g[2]));
  o[0] = fma(F.ax, g[0], dg * F.dx);
  o[1] = fma(F.ay, g[1], dg * F.dy);
  o[2] = dg;
}

// G += A_i Pi_i A_i' as three rank-one terms: ax a_x a_x' + ay a_y a_y' + kap (A_i d)(A_i d)',
// a_x = (1,0,0, 0, rz, -ry), a_y = (0,1,0, -rz, 0, rx), A_i d = (d, r x d).
QPB_HD void add_leg(double (&G)[21], const Face& F, double rx, double ry, double rz) {
  {
    const double a4 = F.ax * rz, a5 = -F.ax * ry;
    G[ix(0, 0)] += F.ax;
    G[ix(4, 0)] += a4;
    G[ix(5, 0)] += a5;
    G[ix(4, 4)] = fma(a4, rz, G[ix(4, 4)]);
    G[ix(5, 4)] = fma(a5, rz, G[ix(5, 4)]);
    G[ix(5, 5)] = fma(-a5, ry, G[ix(5, 5)]);
  }
  {
    const double a3 = -F.ay * rz, a5 = F.ay * rx;
    G[ix(1, 1)] += F.ay;
    G[ix(3, 1)] += a3;
    G[ix(5, 1)] += a5;
    G[ix(3, 3)] = fma(-a3, rz, G[ix(3, 3)]);
    G[ix(5, 3)] = fma(a3, rx, G[ix(5, 3)]);
    G[ix(5, 5)] = fma(a5, rx, G[ix(5, 5)]);
  }
  {
    double v[6], kv[6];
    v[0] = F.dx;
    v[1] = F.dy;
    v[2] = 1.0;
    v[3] = fma(-rz, F.dy, ry);
    v[4] = fma(rz, F.dx, -rx);
    v[5] = fma(rx, F.dy, -ry * F.dx);
#pragma unroll
    for (int i = 0; i < 6; i++) kv[i] = F.kap * v[i];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < 6; j++)
        if (j <= i) G[ix(i, j)] = fma(kv[i], v[j], G[ix(i, j)]);
  }
}

// In-place Cholesky of the packed 6x6 matrix: G <- L (the diagonal holds 1 / L_jj).  False if not positive definite.
QPB_HD bool chol6(double (&G)[21]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double d = G[ix(j, j)];
#pragma unroll
    for (int k = 0; k < 6; k++)
      if (k < j) d = fma(-G[ix(j, k)], G[ix(j, k)], d);
    ok = ok && (d > 0.0) && (d < 1e300);
    const double rs = rsqrt_fast(d);
    G[ix(j, j)] = rs;
#pragma unroll
    for (int i = 0; i < 6; i++)
      if (i > j) {
        double t = G[ix(i, j)];
#pragma unroll
        for (int k = 0; k < 6; k++)
          if (k < j) t = fma(-G[ix(i, k)], G[ix(j, k)], t);
        G[ix(i, j)] = t * rs;
      }
  }
  return ok;
}

// y <- (L L')^-1 y
QPB_HD void chol6_solve(const double (&L)[21], double (&y)[6]) {
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double t = y[j];
#pragma unroll
    for (int k = 0; k < 6; k++)
      if (k < j) t = fma(-L[ix(j, k)], y[k], t);
    y[j] = t * L[ix(j, j)];
  }
#pragma unroll
  for (int jj = 0; jj < 6; jj++) {
    const int j = 5 - jj;
    double t = y[j];
#pragma unroll
    for (int k = 0; k < 6; k++)
      if (k > j) t = fma(-L[ix(k, j)], y[k], t);
    y[j] = t * L[ix(j, j)];
  }
}

// g = A_i' y = y_lin - r x y_ang
QPB_HD void At_y(const double (&y)[6], double rx, double ry, double rz, double (&g)[3]) {
  g[0] = y[0] - (ry * y[5] - rz * y[4]);
  g[1] = y[1] - (rz * y[3] - rx * y[5]);
  g[2] = y[2] - (rx * y[4] - ry * y[3]);
}

// Solver state of one QP (registers on the device).
struct State {
  double f[12];   // world-frame forces
  double u[12];   // multiplier of the active row of group g of leg i at 3 i + g
  double sg[12];  // its sign (0 = group inactive)
  double r[12];   // lever arms R p_i
  uint32_t stance;  // bit i = leg i in contact
  int p;          // row being added (3 leg + group), -1 = none pending
  double ps;      // its sign
  double up;      // its multiplier so far
  int iters, status;
  bool done;
};

// Multipliers of the active rows of one leg from h = N u:  (sx ux, sy uy, mu (ux + uy) + sz uz) = h
QPB_HD void leg_multipliers(const FastParams& K, const double* sg, const double (&h)[3], double (&o)[3]) {
  o[0] = sg[0] * h[0];
  o[1] = sg[1] * h[1];
  o[2] = sg[2] * (h[2] - K.mu * (o[0] + o[1]));
}

// Minimiser on the faces st.sg: fills st.f and st.u.  False when the 6x6 system is not positive definite.
QPB_HD bool face_solve(const FastParams& K, State& st, const double (&b6)[6]) {
  double G[21];
#pragma unroll
  for (int i = 0; i < 21; i++) G[i] = K.wSinv[i];
  Face F[4];
  double pt[12];
  double rhs[6];
#pragma unroll
  for (int i = 0; i < 6; i++) rhs[i] = -b6[i];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool stance = (st.stance >> i) & 1u;
    F[i] = face_of(st.sg[3 * i], st.sg[3 * i + 1], st.sg[3 * i + 2], stance, K);
    add_leg(G, F[i], st.r[3 * i], st.r[3 * i + 1], st.r[3 * i + 2]);
    // a point of the face, then its component normal to the face
    const double sz = st.sg[3 * i + 2];
    const double cz = (sz == 0.0 || !stance) ? 0.0 : (sz > 0.0 ? K.fzmin : K.fzmax);
    const double c[3] = { -st.sg[3 * i] * K.mu * cz, -st.sg[3 * i + 1] * K.mu * cz, cz };
    double pc[3];
    face_proj(F[i], c, pc);
#pragma unroll
    for (int k = 0; k < 3; k++) pt[3 * i + k] = stance ? c[k] - pc[k] : 0.0;
    const double rx = st.r[3 * i], ry = st.r[3 * i + 1], rz = st.r[3 * i + 2];
    rhs[0] += pt[3 * i];
    rhs[1] += pt[3 * i + 1];
    rhs[2] += pt[3 * i + 2];
    rhs[3] += ry * pt[3 * i + 2] - rz * pt[3 * i + 1];
    rhs[4] += rz * pt[3 * i] - rx * pt[3 * i + 2];
    rhs[5] += rx * pt[3 * i + 1] - ry * pt[3 * i];
  }
  const bool ok = chol6(G);
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] = K.w * rhs[i];
  chol6_solve(G, y);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double g[3], pg[3], h[3], m[3];
    At_y(y, st.r[3 * i], st.r[3 * i + 1], st.r[3 * i + 2], g);
    face_proj(F[i], g, pg);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      st.f[3 * i + k] = fma(-K.inv_w, pg[k], pt[3 * i + k]);
      h[k] = 2.0 * fma(K.w, st.f[3 * i + k], g[k]);
    }
    leg_multipliers(K, st.sg + 3 * i, h, m);
#pragma unroll
    for (int k = 0; k < 3; k++) st.u[3 * i + k] = m[k];
  }
  return ok;
}

// Decode a 24-bit working-set word (2 bits per group: 0 none, 1 row A, 2 row B) into signs.
QPB_HD void wset_decode(uint32_t word, uint32_t stance, double (&sg)[12]) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const uint32_t c = (word >> (6 * i + 2 * g)) & 3u;
      const bool st = (stance >> i) & 1u;
      // A rows: -1 on x, y and +1 on z; B rows the opposite
      const double sA = g < 2 ? -1.0 : 1.0;
      sg[3 * i + g] = (!st || c == 0u || c == 3u) ? 0.0 : (c == 1u ? sA : -sA);
    }
}
QPB_HD uint32_t wset_encode(const double (&sg)[12]) {
  uint32_t word = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const double s = sg[3 * i + g];
      const bool isA = g < 2 ? (s < 0.0) : (s > 0.0);
      const uint32_t c = s == 0.0 ? 0u : (isA ? 1u : 2u);
      word |= c << (6 * i + 2 * g);
    }
  return word;
}

// Start: minimiser on the hinted faces if that is a dual-feasible pair, else the unconstrained minimiser.
QPB_HD void start(const FastParams& K, State& st, const double (&b6)[6], uint32_t hint, bool have_hint) {
  st.p = -1;
  st.ps = 0.0;
  st.up = 0.0;
  st.iters = 0;
  st.status = QPB_OK;
  st.done = false;
  bool warm = have_hint && (hint & 0xffffffu) != 0u;
  if (warm) {
    wset_decode(hint, st.stance, st.sg);
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) st.sg[i] = 0.0;
  }
  bool ok = true;
#pragma unroll 1
  for (int attempt = 0; attempt < 2; attempt++) {  // a loop so that face_solve is emitted once
    ok = face_solve(K, st, b6);
    bool feas = true;
#pragma unroll
    for (int i = 0; i < 12; i++) feas = feas && !(st.sg[i] != 0.0 && !(st.u[i] >= 0.0));
    if (!warm || feas) break;
    warm = false;  // the hinted faces are not a dual-feasible pair: cold start
#pragma unroll
    for (int i = 0; i < 12; i++) st.sg[i] = 0.0;
  }
  if (!ok) {
    st.status = QPB_BAD_INPUT;
    st.done = true;
  }
}

// After the loop has ended on a working set: recompute the minimiser on those faces from scratch (one 6x6 solve), so the
// answer carries the rounding of ONE solve instead of the accumulated steps, and re-check it.  The loop's decision was
// taken on the accumulated iterate, so a row may sit within rounding of its bound on the wrong side; anything beyond
// that means the working set is not optimal and the QP is reported unsolved.
QPB_HD void polish(const FastParams& K, State& st, const double (&b6)[6]) {
  if (st.status != QPB_OK) return;
  const bool ok = face_solve(K, st, b6);
  bool good = ok;
  const double zscale = 1.0 + fmax(fabs(K.fzmin), fabs(K.fzmax));
  double umax = 0.0;
#pragma unroll
  for (int i = 0; i < 12; i++) umax = fmax(umax, fabs(st.u[i]));
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if ((st.stance >> i) & 1u) {
      const double fx = st.f[3 * i], fy = st.f[3 * i + 1], fz = st.f[3 * i + 2];
      const double base = K.mu * fz;
      const double loose = -1e-6 * (zscale + fabs(base));
      good = good && (base - fabs(fx) >= loose) && (base - fabs(fy) >= loose) && (fz - K.fzmin >= loose) && (K.fzmax - fz >= loose);
#pragma unroll
      for (int k = 0; k < 3; k++) good = good && (st.sg[3 * i + k] == 0.0 || st.u[3 * i + k] >= -1e-6 * umax);
    }
  }
  if (!good) st.status = QPB_MAX_ITER;
}

// One working-set change of Goldfarb-Idnani (or the optimality test that ends the solve).
QPB_HD void iterate(const FastParams& K, State& st) {
  if (st.done) return;
  // (1) most violated row, if no row is pending
  if (st.p < 0) {
    double best = 0.0;
    int bidx = -1;
    double bsgn = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool stance = (st.stance >> i) & 1u;
      const double fx = st.f[3 * i], fy = st.f[3 * i + 1], fz = st.f[3 * i + 2];
      const double base = K.mu * fz;
      const double slx = base - fabs(fx), sly = base - fabs(fy);
      const double sA = fz - K.fzmin, sB = K.fzmax - fz;
      const double slz = fmin(sA, sB);
      if (stance) {
        if (st.sg[3 * i] == 0.0) {
          if (slx < -1e-9 && slx < best) { best = slx; bidx = 3 * i; bsgn = fx > 0.0 ? -1.0 : 1.0; }
        }  // (the other row of an active group can only be violated while fz < 0: the fz row of this leg goes first)
        if (st.sg[3 * i + 1] == 0.0) {
          if (sly < -1e-9 && sly < best) { best = sly; bidx = 3 * i + 1; bsgn = fy > 0.0 ? -1.0 : 1.0; }
        }
        if (st.sg[3 * i + 2] == 0.0) {
          if (slz < K.ntol_z && slz < best) { best = slz; bidx = 3 * i + 2; bsgn = sA < sB ? 1.0 : -1.0; }
        }
      }
    }
    if (bidx < 0) {  // no inactive group is violated: the loop ends here (polish() re-checks every row)
      st.done = true;
      return;
    }
    st.p = bidx;
    st.ps = bsgn;
    st.up = 0.0;
  }
  if (st.iters >= K.max_iter) {
    st.status = QPB_MAX_ITER;
    st.done = true;
    return;
  }
  st.iters++;
  const int p = st.p;
  const int pl = (p * 11) >> 5, pg = p - 3 * pl;  // leg and group of row p
  const double n[3] = { pg == 0 ? st.ps : 0.0, pg == 1 ? st.ps : 0.0, pg == 2 ? st.ps : K.mu };
  const double dp = pg == 2 ? (st.ps > 0.0 ? K.fzmin : -K.fzmax) : 0.0;  // row p: n' f >= dp

  // (2) faces, G = w S^-1 + sum A_i Pi_i A_i', right-hand side (1/2) A_l Pi_l n
  double G[21];
#pragma unroll
  for (int i = 0; i < 21; i++) G[i] = K.wSinv[i];
  Face F[4];
  double tn[3] = { 0.0, 0.0, 0.0 };
  double yh[6] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
  double sp = -dp;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool stance = (st.stance >> i) & 1u;
    F[i] = face_of(st.sg[3 * i], st.sg[3 * i + 1], st.sg[3 * i + 2], stance, K);
    const double rx = st.r[3 * i], ry = st.r[3 * i + 1], rz = st.r[3 * i + 2];
    add_leg(G, F[i], rx, ry, rz);
    if (i == pl) {
      face_proj(F[i], n, tn);
      yh[0] = 0.5 * tn[0];
      yh[1] = 0.5 * tn[1];
      yh[2] = 0.5 * tn[2];
      yh[3] = 0.5 * (ry * tn[2] - rz * tn[1]);
      yh[4] = 0.5 * (rz * tn[0] - rx * tn[2]);
      yh[5] = 0.5 * (rx * tn[1] - ry * tn[0]);
      sp += n[0] * st.f[3 * i] + n[1] * st.f[3 * i + 1] + n[2] * st.f[3 * i + 2];
    }
  }
  const bool pd = chol6(G);
  chol6_solve(G, yh);

  // (3) step direction z, its curvature zeta = n' z, multiplier directions r, blocking row
  double z[12], rr[12];
  double zeta = 0.0;
  double ub = 1.0, rb = 0.0;  // best ratio so far as a fraction ub / rb (rb = 0: none)
  int kb = -1;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double g[3], pgv[3], h[3], m[3];
    At_y(yh, st.r[3 * i], st.r[3 * i + 1], st.r[3 * i + 2], g);
    face_proj(F[i], g, pgv);
    const double ci = (i == pl) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      z[3 * i + k] = K.inv_w * fma(0.5 * ci, tn[k], -pgv[k]);
      h[k] = fma(ci, n[k], -2.0 * fma(K.w, z[3 * i + k], g[k]));
    }
    zeta += ci * (n[0] * z[3 * i] + n[1] * z[3 * i + 1] + n[2] * z[3 * i + 2]);
    leg_multipliers(K, st.sg + 3 * i, h, m);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      rr[3 * i + k] = m[k];
      const double uu = st.u[3 * i + k] > 0.0 ? st.u[3 * i + k] : 0.0;  // rounding can leave -1e-17
      const bool cand = st.sg[3 * i + k] != 0.0 && m[k] > 0.0;
      if (cand && uu * rb < ub * m[k]) {  // uu / m[k] < ub / rb
        ub = uu;
        rb = m[k];
        kb = 3 * i + k;
      }
    }
  }
  if (!pd || !(zeta > 0.0)) {  // cannot happen within the face family (G is positive definite, n is independent)
    st.status = QPB_MAX_ITER;
    st.done = true;
    return;
  }
  const bool has1 = kb >= 0;
  const double t1 = has1 ? ub * rcp_fast(rb) : 0.0;
  const double t2 = -sp * rcp_fast(zeta);
  const bool full = !has1 || t2 <= t1;
  const double t = full ? t2 : t1;
  // (4) step
#pragma unroll
  for (int i = 0; i < 12; i++) {
    st.f[i] = fma(t, z[i], st.f[i]);
    st.u[i] = fma(-t, rr[i], st.u[i]);
  }
  st.up += t;
  // (5) working-set change: row p enters (full step) or the blocking row leaves (partial step)
  const int idx = full ? p : kb;
  const double sv = full ? st.ps : 0.0, uv = full ? st.up : 0.0;
#pragma unroll
  for (int i = 0; i < 12; i++)
    if (i == idx) {
      st.sg[i] = sv;
      st.u[i] = uv;
    }
  if (full) st.p = -1;
}

// ---- whole-record wrappers: what one thread does before and after the iteration loop ----------------------------

QPB_HD uint32_t stance_mask(uint32_t cbytes) {
  return ((cbytes & 0xffu) ? 1u : 0u) | ((cbytes & 0xff00u) ? 2u : 0u) | ((cbytes & 0xff0000u) ? 4u : 0u) |
         ((cbytes & 0xff000000u) ? 8u : 0u);
}

// rec: slots 0..47 of the state record (attitudes, twists, feet).  hint: bit 31 set = bits 0..23 hold a working set.
template <class Params>
QPB_HD void setup(const Params& P, const FastParams& K, const double* rec, uint32_t cbytes, uint32_t hint, State& st,
                  double (&b6)[6]) {
  st.stance = stance_mask(cbytes);
  bool fin = true;
#pragma unroll
  for (int i = 0; i < 48; i++) fin = fin && (fabs(rec[i]) <= 1.79769313486231570e308);  // false for NaN and +-inf
  pd_rhs(P, rec, b6);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 3; k++)  // lever arms r_i = R p_i, balance_controller.cpp:245-248
      st.r[3 * i + k] = rec[kR + 3 * k] * rec[kFeet + 3 * i] + rec[kR + 3 * k + 1] * rec[kFeet + 3 * i + 1] +
                        rec[kR + 3 * k + 2] * rec[kFeet + 3 * i + 2];
#pragma unroll
  for (int i = 0; i < 6; i++) fin = fin && (fabs(b6[i]) <= 1.79769313486231570e308);
  if (!fin) {  // keep the arithmetic finite; the QP is reported as bad input
#pragma unroll
    for (int i = 0; i < 12; i++) st.r[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) b6[i] = 0.0;
  }
  start(K, st, b6, hint, (hint >> 31) != 0u);
  if (!fin) {
    st.status = QPB_BAD_INPUT;
    st.done = true;
  }
}

// World-frame solution -> body-frame GRF (balance_controller.cpp:218-232) and tau = J^T f (kinematics.cpp:162-188,
// 218-231, clamp commander_node.cpp:526).  R: Rwb (9), q: joint angles (12).
template <class Params>
QPB_HD void finish(const Params& P, const double* R, const double* q, const State& st, double (&grf)[12], double (&tau)[12]) {
  const bool good = st.status == QPB_OK;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool on = good && ((st.stance >> i) & 1u);
    const double f0 = st.f[3 * i], f1 = st.f[3 * i + 1], f2 = st.f[3 * i + 2];
    double fb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) fb[k] = on ? -1.0 * (R[k] * f0 + R[3 + k] * f1 + R[6 + k] * f2) : 0.0;
    double s1, c1, s2, c2, s23, c23;
    const bool qok = st.status != QPB_BAD_INPUT;
    sincos(qok ? q[3 * i] : 0.0, &s1, &c1);
    sincos(qok ? q[3 * i + 1] : 0.0, &s2, &c2);
    sincos(qok ? q[3 * i + 1] + q[3 * i + 2] : 0.0, &s23, &c23);
    double t[3];
    leg_jt(P.link[3 * i], P.link[3 * i + 1], P.link[3 * i + 2], s1, c1, s2, c2, s23, c23, fb[0], fb[1], fb[2], t);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double v = t[k];
      if (P.clamp_tau) v = fmin(fmax(v, P.tau_min), P.tau_max);
      grf[3 * i + k] = fb[k];
      tau[3 * i + k] = on ? v : 0.0;
    }
  }
}

// Host-side derivation of FastParams.  Returns false when W is not a multiple of the identity (the general kernel
// must be used) or S cannot be inverted.
inline bool make_fast_params(const qpb_params& P, FastParams& K) {
  const double w = P.W[0];
  for (int i = 0; i < 12; i++)
    for (int j = 0; j < 12; j++)
      if (P.W[12 * i + j] != (i == j ? w : 0.0)) return false;
  if (!(w > 0.0)) return false;
  // S^-1 by Gauss-Jordan on the SPD matrix (6x6, host, once per controller)
  double a[6][12];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      a[i][j] = 0.5 * (P.S[6 * i + j] + P.S[6 * j + i]);
      a[i][6 + j] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < 6; c++) {
    int piv = c;
    for (int i = c + 1; i < 6; i++)
      if (fabs(a[i][c]) > fabs(a[piv][c])) piv = i;
    if (!(fabs(a[piv][c]) > 0.0)) return false;
    if (piv != c)
      for (int j = 0; j < 12; j++) { const double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
    const double inv = 1.0 / a[c][c];
    for (int j = 0; j < 12; j++) a[c][j] *= inv;
    for (int i = 0; i < 6; i++)
      if (i != c) {
        const double m = a[i][c];
        for (int j = 0; j < 12; j++) a[i][j] -= m * a[c][j];
      }
  }
  K.w = w;
  K.inv_w = 1.0 / w;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) K.wSinv[ix(i, j)] = w * 0.5 * (a[i][6 + j] + a[j][6 + i]);
  K.mu = P.mu;
  K.k1 = 1.0 / (1.0 + P.mu * P.mu);
  K.k2 = 1.0 / (1.0 + 2.0 * P.mu * P.mu);
  K.fzmin = P.fzmin;
  K.fzmax = P.fzmax;
  K.ntol_z = -1e-9 * (1.0 + fmax(fabs(P.fzmin), fabs(P.fzmax)));
  K.max_iter = P.max_iter;
  return true;
}

}  // namespace tpq
}  // namespace qpb
