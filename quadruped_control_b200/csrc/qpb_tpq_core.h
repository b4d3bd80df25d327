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
// comes from a fresh 6x6 Cholesky factorisation of G = w S^-1 + sum_i A_i Pi_i A_i'; G itself follows the working
// set by one rank-one update per change (G -+= v v' / (n' t), v = A_l t, t = the row's normal projected on the face
// without that row).  State per QP: f (12), multipliers (12), lever arms (12), G (21), a 24-bit working-set word.
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

#include <string.h>

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

// bit patterns of doubles (ordering tricks below): host and device
QPB_HD int hi32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  return (int)(b >> 32);
#endif
}
QPB_HD double flip_sign(double x, bool flip) {  // flip ? -x : x, on the integer datapath
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(x) ^ (flip ? (int)0x80000000 : 0), __double2loint(x));
#else
  return flip ? -x : x;
#endif
}

// Working set: one 24-bit word, 2 bits per (leg, group): 0 none, 1 row A, 2 row B, at bit 6 leg + 2 group.
//   x rows (s, 0, mu), y rows (0, s, mu): A has s = -1 (-f + mu fz >= 0), B has s = +1 (f + mu fz >= 0)
//   z rows (0, 0, s): A has s = +1 (fz >= fzmin), B has s = -1 (-fz >= -fzmax)
// Face of one leg: tangent-space projector Pi = diag(!X, !Y, 0) + kap d d', d = (dx, dy, 1)  (X: an x row is active, ...).
// A swing leg is pinned at zero: Pi = 0 and no rows.
struct Leg {
  bool X, Y;          // x / y component fixed by the face (or swing leg)
  uint32_t cx, cy, cz;  // row codes
  double dx, dy, kap;
};

// code 1 -> +mu, code 2 -> -mu, 0 -> 0  (= -s mu for the row's own coefficient s), on the integer datapath
QPB_HD double signed_mu(const FastParams& K, uint32_t c) {
#if defined(__CUDA_ARCH__)
  const int m = c ? -1 : 0;
  return __hiloint2double((__double2hiint(K.mu) & m) ^ (int)((c & 2u) << 30), __double2loint(K.mu) & m);
#else
  return c == 1u ? K.mu : (c == 2u ? -K.mu : 0.0);
#endif
}

QPB_HD Leg leg_of(uint32_t codes /* word >> 6 leg */, bool stance, const FastParams& K) {
  Leg L;
  const uint32_t c = stance ? codes : 0u;
  L.cx = c & 3u;
  L.cy = (c >> 2) & 3u;
  L.cz = (c >> 4) & 3u;
  const bool ax = L.cx != 0u, ay = L.cy != 0u;
  L.X = ax || !stance;
  L.Y = ay || !stance;
  L.dx = signed_mu(K, L.cx);
  L.dy = signed_mu(K, L.cy);
  const double k12 = (ax && ay) ? K.k2 : K.k1;
  const double k01 = (ax || ay) ? k12 : 1.0;
  L.kap = (L.cz != 0u || !stance) ? 0.0 : k01;
  return L;
}

QPB_HD void leg_proj(const Leg& L, const double (&g)[3], double (&o)[3]) {
  const double dg = L.kap * fma(L.dx, g[0], fma(L.dy, g[1], g[2]));
  o[0] = L.X ? dg * L.dx : g[0];  // x free => dx = 0 => (Pi g)_x = g_x
  o[1] = L.Y ? dg * L.dy : g[1];
  o[2] = dg;
}

// normal of row (group g, code c) restricted to its leg, and its bound: n' f >= dp
QPB_HD void row_normal(const FastParams& K, int g, uint32_t c, double (&n)[3], double& dp) {
  const double s = (g < 2) == (c == 1u) ? -1.0 : 1.0;
  n[0] = g == 0 ? s : 0.0;
  n[1] = g == 1 ? s : 0.0;
  n[2] = g == 2 ? s : K.mu;
  dp = g == 2 ? (c == 1u ? K.fzmin : -K.fzmax) : 0.0;
}

// Multipliers of the active rows of one leg from h = N u:  (sx ux, sy uy, mu (ux + uy) + sz uz) = h
QPB_HD void leg_multipliers(const FastParams& K, const Leg& L, const double (&h)[3], double (&o)[3]) {
  o[0] = L.cx == 0u ? 0.0 : flip_sign(h[0], L.cx == 1u);
  o[1] = L.cy == 0u ? 0.0 : flip_sign(h[1], L.cy == 1u);
  const double t = h[2] - K.mu * (o[0] + o[1]);
  o[2] = L.cz == 0u ? 0.0 : flip_sign(t, L.cz == 2u);
}

// G += c v v'  (packed lower triangle)
QPB_HD void rank1(double (&G)[21], const double (&v)[6], double c) {
  double cv[6];
#pragma unroll
  for (int i = 0; i < 6; i++) cv[i] = c * v[i];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++)
      if (j <= i) G[ix(i, j)] = fma(cv[i], v[j], G[ix(i, j)]);
}

// G += A_i Pi_i A_i' as three rank-one terms: !X a_x a_x' + !Y a_y a_y' + kap (A_i d)(A_i d)',
// a_x = (1,0,0, 0, rz, -ry), a_y = (0,1,0, -rz, 0, rx), A_i d = (d, r x d).
QPB_HD void add_leg(double (&G)[21], const Leg& L, double rx, double ry, double rz) {
  const double fx = L.X ? 0.0 : 1.0, fy = L.Y ? 0.0 : 1.0;
  {
    const double a4 = fx * rz, a5 = -fx * ry;
    G[ix(0, 0)] += fx;
    G[ix(4, 0)] += a4;
    G[ix(5, 0)] += a5;
    G[ix(4, 4)] = fma(a4, rz, G[ix(4, 4)]);
    G[ix(5, 4)] = fma(a5, rz, G[ix(5, 4)]);
    G[ix(5, 5)] = fma(-a5, ry, G[ix(5, 5)]);
  }
  {
    const double a3 = -fy * rz, a5 = fy * rx;
    G[ix(1, 1)] += fy;
    G[ix(3, 1)] += a3;
    G[ix(5, 1)] += a5;
    G[ix(3, 3)] = fma(-a3, rz, G[ix(3, 3)]);
    G[ix(5, 3)] = fma(a3, rx, G[ix(5, 3)]);
    G[ix(5, 5)] = fma(a5, rx, G[ix(5, 5)]);
  }
  const double v[6] = { L.dx, L.dy, 1.0, fma(-rz, L.dy, ry), fma(rz, L.dx, -rx), fma(rx, L.dy, -ry * L.dx) };
  rank1(G, v, L.kap);
}

// In-place Cholesky of the packed 6x6 matrix: G <- L (the diagonal holds 1 / L_jj).  False if not positive definite.
// (Loops have constant trip counts with guards: the front end then keeps G in registers.)  LOOP: the factorisation of
// the active-set loop, which only takes decisions with it (the polish recomputes the answer): a 2^-40 rsqrt shortens the
// six-deep chain of dependent operations that bounds an iteration.
template <bool LOOP = false>
QPB_HD bool chol6(double (&G)[21]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double d = G[ix(j, j)];
#pragma unroll
    for (int k = 0; k < 6; k++)
      if (k < j) d = fma(-G[ix(j, k)], G[ix(j, k)], d);
    ok = ok && (d > 0.0) && (d < 1e300);
    const double rs = LOOP ? rsqrt_loop(d) : rsqrt_fast(d);
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
// v = A_i t = (t, r x t)
QPB_HD void A_t(const double (&t)[3], double rx, double ry, double rz, double (&v)[6]) {
  v[0] = t[0];
  v[1] = t[1];
  v[2] = t[2];
  v[3] = ry * t[2] - rz * t[1];
  v[4] = rz * t[0] - rx * t[2];
  v[5] = rx * t[1] - ry * t[0];
}

// Solver state of one QP (registers on the device).  The 6x6 matrix G = w S^-1 + sum_i A_i Pi_i A_i' of the current
// working set lives outside (per-lane shared memory on the device): it is loaded, factorised and rank-one updated
// once per working-set change.
struct State {
  double f[12];   // world-frame forces
  double u[12];   // multiplier of the active row of group g of leg i at 3 i + g
  double r[12];   // lever arms R p_i
  uint32_t word;  // working set (24 bits)
  uint32_t stance;  // bit i = leg i in contact
  int p;          // row being added: 3 leg + group, -1 = none pending
  uint32_t pc;    // its code (1 = A, 2 = B)
  double up;      // its multiplier so far
  int iters, status;
  bool done;
};

// Minimiser on the faces st.word: fills st.f, st.u and G (unfactorised).  False when G is not positive definite.
QPB_HD bool face_solve(const FastParams& K, State& st, const double (&b6)[6], double (&G)[21]) {
#pragma unroll
  for (int i = 0; i < 21; i++) G[i] = K.wSinv[i];
  Leg L[4];
  double pt[12];
  double rhs[6];
#pragma unroll
  for (int i = 0; i < 6; i++) rhs[i] = -b6[i];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool stance = (st.stance >> i) & 1u;
    L[i] = leg_of(st.word >> (6 * i), stance, K);
    const double rx = st.r[3 * i], ry = st.r[3 * i + 1], rz = st.r[3 * i + 2];
    add_leg(G, L[i], rx, ry, rz);
    // a point of the face (fz on its bound if a z row is active, x / y on the pyramid sides), then its component
    // normal to the face
    const double cz = L[i].cz == 0u ? 0.0 : (L[i].cz == 1u ? K.fzmin : K.fzmax);
    const double c[3] = { L[i].dx * cz, L[i].dy * cz, cz };
    double pc[3], v[6];
    leg_proj(L[i], c, pc);
#pragma unroll
    for (int k = 0; k < 3; k++) pc[k] = c[k] - pc[k];
    A_t(pc, rx, ry, rz, v);
#pragma unroll
    for (int k = 0; k < 3; k++) pt[3 * i + k] = pc[k];
#pragma unroll
    for (int k = 0; k < 6; k++) rhs[k] += v[k];
  }
  double Lc[21];
#pragma unroll
  for (int i = 0; i < 21; i++) Lc[i] = G[i];
  const bool ok = chol6(Lc);
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] = K.w * rhs[i];
  chol6_solve(Lc, y);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double g[3], pg[3], h[3], m[3];
    At_y(y, st.r[3 * i], st.r[3 * i + 1], st.r[3 * i + 2], g);
    leg_proj(L[i], g, pg);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      st.f[3 * i + k] = fma(-K.inv_w, pg[k], pt[3 * i + k]);
      h[k] = 2.0 * fma(K.w, pt[3 * i + k], g[k] - pg[k]);  // = 2 (g + w f): the component normal to the face
    }
    leg_multipliers(K, L[i], h, m);
#pragma unroll
    for (int k = 0; k < 3; k++) st.u[3 * i + k] = m[k];
  }
  return ok;
}

// keep only well-formed codes of stance legs (a hint is untrusted input)
QPB_HD uint32_t wset_sanitize(uint32_t word, uint32_t stance) {
  const uint32_t both = word & (word >> 1) & 0x555555u;  // groups whose code is 3 (both rows at once): not a face
  uint32_t legs = 0u;
#pragma unroll
  for (int i = 0; i < 4; i++) legs |= ((stance >> i) & 1u) ? (63u << (6 * i)) : 0u;
  return word & ~(both | (both << 1)) & legs;
}

// Selection key of one leg: the most violated row among its groups that are not active, 0 if none.  The high word of
// (slack - tolerance) orders negative doubles by magnitude as an unsigned integer; low 5 bits = 3 leg + group, + 16 for
// row B.  (The other row of an ACTIVE group can only be violated while fz < 0; the fz row of that leg then goes first.)
QPB_HD uint32_t leg_key(const FastParams& K, double fx, double fy, double fz, uint32_t codes, bool stance, int i) {
  const double base = fma(K.mu, fz, 1e-9);
  const double slx = base - fabs(fx), sly = base - fabs(fy);
  const double sA = (fz - K.fzmin) - K.ntol_z, sB = (K.fzmax - fz) - K.ntol_z;
  const uint32_t id = (uint32_t)(3 * i);
  const uint32_t kx = ((uint32_t)hi32(slx) & ~31u) | id | (hi32(fx) < 0 ? 16u : 0u);  // fx > 0: row A binds
  const uint32_t ky = ((uint32_t)hi32(sly) & ~31u) | (id + 1u) | (hi32(fy) < 0 ? 16u : 0u);
  const uint32_t ka = ((uint32_t)hi32(sA) & ~31u) | (id + 2u);
  const uint32_t kb = ((uint32_t)hi32(sB) & ~31u) | (id + 2u) | 16u;
  const uint32_t mx = (stance && (codes & 3u) == 0u) ? kx : 0u;
  const uint32_t my = (stance && (codes & 12u) == 0u) ? ky : 0u;
  const uint32_t mz = (stance && (codes & 48u) == 0u) ? (ka > kb ? ka : kb) : 0u;
  const uint32_t m1 = mx > my ? mx : my;
  const uint32_t m2 = m1 > mz ? m1 : mz;
  return (m2 >> 31) ? m2 : 0u;  // only violated rows (negative slack) count
}

// Rows violated at st.f among the groups that are not active, at most one per group (the binding side), as a
// working-set word.  Same tolerances as the loop's selection.
QPB_HD uint32_t violated_rows(const FastParams& K, const State& st) {
  uint32_t w = 0u;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t c = st.word >> (6 * i);
    const double fx = st.f[3 * i], fy = st.f[3 * i + 1], fz = st.f[3 * i + 2];
    const double base = fma(K.mu, fz, 1e-9);
    uint32_t add = 0u;
    if ((c & 3u) == 0u && base - fabs(fx) < 0.0) add |= hi32(fx) < 0 ? 2u : 1u;  // fx > 0: row A binds
    if ((c & 12u) == 0u && base - fabs(fy) < 0.0) add |= hi32(fy) < 0 ? 8u : 4u;
    if ((c & 48u) == 0u) {
      if ((fz - K.fzmin) - K.ntol_z < 0.0) add |= 16u;
      else if ((K.fzmax - fz) - K.ntol_z < 0.0) add |= 32u;
    }
    if ((st.stance >> i) & 1u) w |= add << (6 * i);
  }
  return w;
}

#ifndef QPB_START_SAT_MIN
#define QPB_START_SAT_MIN 4  // violated rows from which a QP counts as heavily loaded
#endif
QPB_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
// A guess of the optimal working set for heavily loaded QPs.  At the optimum of a hard-pushed robot most groups are
// active (10 of 12 on BASELINE config 2), far more than the rows the unconstrained minimiser violates.  So when at least
// QPB_START_SAT_MIN rows are violated, the block round adds, beside them, a row in every free x / y group of the
// stance legs -- the side the force points to (mode 1; mode 2 adds the nearer fz bound as well, modes 3 / 4 only touch
// legs that have a violated row: all measured worse).  Rows guessed wrong come out with a negative multiplier and the
// drop rounds remove them; the result only has to be a dual-feasible pair.  Working-set changes left to the loop on
// config 2 / 3 / stress: 5.7 -> 3.7, 2.7 -> 2.0, 6.0 -> 4.1 per QP; light profile 2.6 -> 3.0 (lightly loaded QPs rarely
// reach the threshold).  Throughput +4 % / +4 % / +6 % / -0.5 % (profiles/r02_saturate_ab.txt): the mean falls, the
// longest QPs (25-30 changes) that set the length of the loop pass do not.
QPB_HD uint32_t saturated_rows(const FastParams& K, const State& st, uint32_t viol) {
  uint32_t w = viol;
  const bool all_legs = popc32(viol) >= QPB_START_SAT_MIN;
  (void)all_legs;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (!((st.stance >> i) & 1u)) continue;
#if QPB_START_SATURATE == 3
    if (((viol >> (6 * i)) & 63u) == 0u) continue;  // only legs that have a violated row
#elif QPB_START_SATURATE == 4
    if (!all_legs && ((viol >> (6 * i)) & 63u) == 0u) continue;
#endif
    const uint32_t c = (st.word | viol) >> (6 * i);
    const double fx = st.f[3 * i], fy = st.f[3 * i + 1], fz = st.f[3 * i + 2];
    (void)fz;  // (mode 2 only)
    uint32_t add = 0u;
    if ((c & 3u) == 0u) add |= hi32(fx) < 0 ? 2u : 1u;
    if ((c & 12u) == 0u) add |= hi32(fy) < 0 ? 8u : 4u;
#if QPB_START_SATURATE >= 2
    if ((c & 48u) == 0u) add |= (fz - K.fzmin < K.fzmax - fz) ? 16u : 32u;
#endif
    w |= add << (6 * i);
  }
  return w;
}

#ifndef QPB_START_SATURATE
#define QPB_START_SATURATE 1  // the block round of a heavily loaded QP also guesses the rows of its free x / y groups (0: violated rows only)
#endif
#ifndef QPB_START_ADDS
#define QPB_START_ADDS 1
#endif
#ifndef QPB_START_DROPS
#define QPB_START_DROPS 2
#endif
constexpr int kStartAdds = QPB_START_ADDS;    // block rounds that add every violated row at once
constexpr int kStartDrops = QPB_START_DROPS;  // rounds that drop every row with a negative multiplier, per add round
constexpr int kStartSolves = 1 + kStartAdds * (1 + kStartDrops) + 1 + kStartDrops;  // bound on the 6x6 solves of start()

// Start of the active-set method.  Goldfarb-Idnani may start from ANY working set whose face minimiser has non-negative
// multipliers (a dual-feasible pair); the closer that set is to the optimal one, the fewer one-row-at-a-time changes
// the loop needs.  So, before the loop: solve on the hinted faces (warm start = the reference's hotstart) or on none
// (the unconstrained minimiser); then up to kStartAdds block rounds: add EVERY violated row at once, re-solve, drop
// every row whose multiplier came out negative (up to kStartDrops times), and keep the result only if it is a
// dual-feasible pair.  Every pair that qualifies is handed to commit(st, G, key) -- key = the first row the loop would
// add there, 0 if the pair is already optimal; the last one committed is where the loop
// starts (often it is already optimal: with one block round the loop's working-set changes fall from 12.5 to 5.7 per
// QP on BASELINE config 2 and from 7.0 to 2.7 on config 3).  A solve from scratch costs about as much as one loop
// iteration, so only the first block round -- which stands in for ~6 iterations -- pays; measured on the B200, one
// round beats none, two and three (profiles/r02_start_budget.txt).
template <class Commit>
QPB_HD void start(const FastParams& K, State& st, const double (&b6)[6], uint32_t hint, bool have_hint, double (&G)[21],
                  Commit& commit) {
  st.p = -1;
  st.pc = 0u;
  st.up = 0.0;
  st.iters = 0;
  st.status = QPB_OK;
  st.done = false;
  uint32_t word = have_hint ? wset_sanitize(hint, st.stance) : 0u;
  bool committed = false, ok = true;
  int adds = 0, drops = 0, rounds = 0;
#pragma unroll 1
  for (int pass = 0; pass < kStartSolves; pass++) {  // a loop so that face_solve is emitted once
    st.word = word;
    st.iters = rounds;
    ok = face_solve(K, st, b6, G);
    if (!ok) break;
    uint32_t neg = 0u;
#pragma unroll
    for (int i = 0; i < 12; i++)
      if (st.u[i] < 0.0) neg |= 3u << (2 * i);  // inactive groups carry u = 0
    if (neg == 0u) {
      uint32_t key = 0u;  // the row the loop would add first at this pair; 0: nothing is violated, the pair is optimal
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t k = leg_key(K, st.f[3 * i], st.f[3 * i + 1], st.f[3 * i + 2], word >> (6 * i), (st.stance >> i) & 1u, i);
        key = k > key ? k : key;
      }
      commit(st, G, key);
      committed = true;
      uint32_t viol = key ? violated_rows(K, st) : 0u;  // (key = 0: no slack is negative, the word would be empty)
#if QPB_START_SATURATE
      if (adds == 0 && (QPB_START_SATURATE == 4 ? viol != 0u : popc32(viol) >= QPB_START_SAT_MIN)) viol = saturated_rows(K, st, viol);
#endif
      if (viol == 0u || adds >= kStartAdds || rounds >= K.max_iter) break;  // optimal already / budget spent
      word |= viol;
      adds++;
      drops = 0;
    } else if (drops < kStartDrops && rounds < K.max_iter) {
      word &= ~neg;
      drops++;
    } else if (!committed) {
      word = 0u;  // a hint that does not lead to a dual-feasible pair: cold start (the empty set always qualifies)
      drops = 0;
      rounds = -1;
    } else {
      break;
    }
    rounds++;
  }
  if (!ok || !committed) {
    st.status = QPB_BAD_INPUT;
    st.done = true;
    commit(st, G, 0u);
  }
}

// After the loop has ended on a working set: recompute the minimiser on those faces from scratch (one 6x6 solve), so the
// answer carries the rounding of ONE solve instead of the accumulated steps, and re-check it.  The loop's decision was
// taken on the accumulated iterate, so a row may sit within rounding of its bound on the wrong side; anything beyond
// that means the working set is not optimal and the QP is reported unsolved.
QPB_HD void polish(const FastParams& K, State& st, const double (&b6)[6]) {
  if (st.status != QPB_OK) return;
  double G[21];
  const bool ok = face_solve(K, st, b6, G);
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
      for (int k = 0; k < 3; k++) good = good && (st.u[3 * i + k] >= -1e-6 * umax);
    }
  }
  if (!good) st.status = QPB_MAX_ITER;
}

// ---- the iteration loop, LPQ lanes per QP ---------------------------------------------------------------------------
// A QP is iterated on by LPQ = 1, 2 or 4 lanes; each owns LPL = 4 / LPQ legs (their f and u).  Everything a lane needs of
// the other legs is either replicated (lever arms, working-set word, pending row, G in shared memory) or crosses the
// group at three exchange points per working-set change: the most violated row (one integer max), its slack (one sum),
// the blocking row (one fraction min).  The 6x6 factorisation is redundant across the lanes of a QP; the per-leg work
// is not.  With LPQ = 1 this is one thread per QP and the exchanges are the identity.

template <int LPL>
struct Lane {
  double f[3 * LPL], u[3 * LPL];  // own legs: forces and multipliers (row of group g of own leg li at 3 li + g)
  double r[3 * LPL];              // own legs: lever arms (all twelve also sit in the QP's side block, for the row's leg)
  uint32_t word, stance;          // working set, contact mask (replicated, as is everything below)
  int p;                          // row being added: 3 leg + group, -1 = none pending
  uint32_t pc;                    // its code (1 = A, 2 = B)
  double up, sp;                  // its multiplier so far, its slack n' f - bound (< 0 while pending)
  int iters, status;
  bool done;
};

// One lane's share of a solver state.  j = index of the lane within its QP.
template <int LPL>
QPB_HD void lane_init(Lane<LPL>& ln, int j, const double* f12, const double* r12, const double* u12, uint32_t word, uint32_t stance,
                      int status, int iters, uint32_t key) {
#pragma unroll
  for (int i = 0; i < 3 * LPL; i++) {
    ln.f[i] = f12[3 * LPL * j + i];
    ln.u[i] = u12[3 * LPL * j + i];
    ln.r[i] = r12[3 * LPL * j + i];
  }
  ln.word = word;
  ln.stance = stance;
  // the loop enters with its first row chosen (by the set-up pass); its slack follows from row_slack_share()
  ln.p = (key >> 31) ? (int)(key & 15u) : -1;
  ln.pc = (key & 16u) ? 2u : 1u;
  ln.up = 0.0;
  ln.sp = 0.0;
  ln.iters = iters;  // working-set changes already spent by the set-up's block rounds
  ln.status = status;
  ln.done = status != QPB_OK || ln.p < 0;
}

// This lane's share of the slack n' f - bound of the pending row (the lane that owns its leg has all of it).
template <int LPL>
QPB_HD double row_slack_share(const FastParams& K, const Lane<LPL>& ln, int j) {
  const int p = ln.p < 0 ? 0 : ln.p;
  const int pl = (p * 11) >> 5, pg = p - 3 * pl;
  double n[3], dp, share = 0.0;
  row_normal(K, pg, ln.pc, n, dp);
#pragma unroll
  for (int li = 0; li < LPL; li++)
    if (LPL * j + li == pl) share = (n[0] * ln.f[3 * li] + n[1] * ln.f[3 * li + 1] + n[2] * ln.f[3 * li + 2]) - dp;
  return ln.p < 0 ? 0.0 : share;
}

// side block of a QP while it is iterated on (shared memory on the device): G (21), lever arms of all legs (12)
enum : int { kSideG = 0, kSideR = 21, kSideSize = 33 };

// Prepared record: what the set-up pass hands to the loop (64 doubles): r (12), b (6), f (12), u (12), G (21), meta.
// The finishing pass reads only r and b -- the first 144 bytes -- and the QP's result word from a separate array.
enum : int { kPrepR = 0, kPrepB = 12, kPrepF = 18, kPrepU = 30, kPrepG = 42, kPrepMeta = 63, kPrepSize = 64 };

// Exchange 1 (integer max over the group): the most violated row among the groups that are not active (leg_key).
// 0 from lanes that are not selecting.
template <int LPL>
QPB_HD uint32_t select_local(const FastParams& K, const Lane<LPL>& ln, int j) {
  uint32_t best = 0u;
#pragma unroll
  for (int li = 0; li < LPL; li++) {
    const int i = LPL * j + li;  // leg
    const uint32_t k = leg_key(K, ln.f[3 * li], ln.f[3 * li + 1], ln.f[3 * li + 2], ln.word >> (6 * i), (ln.stance >> i) & 1u, i);
    best = best > k ? best : k;
  }
  return (ln.done || ln.p >= 0) ? 0u : best;
}

// After exchange 1: take the row (or finish: nothing is violated).  Returns this lane's contribution to exchange 2
// (a sum over the group): the exact slack of the new row, from the lane that owns its leg.
template <int LPL>
QPB_HD double select_commit(const FastParams& K, Lane<LPL>& ln, int j, uint32_t best, bool& fresh) {
  fresh = false;
  if (!ln.done && ln.p < 0) {
    if ((best >> 31) == 0u) {
      ln.done = true;  // no inactive group is violated: the loop ends here (polish() re-checks every row)
    } else {
      fresh = true;
      ln.p = (int)(best & 15u);
      ln.pc = (best & 16u) ? 2u : 1u;
      ln.up = 0.0;
    }
  }
  return fresh ? row_slack_share<LPL>(K, ln, j) : 0.0;
}

template <int LPL>
struct StepTmp {
  double z[3 * LPL], rr[3 * LPL];
  double a[6];
  double zeta0, zeta;
  bool act;
};

// Between exchanges 2 and 3: step direction on the own legs, multiplier directions, and the own candidate for the
// blocking row as a fraction ub / rb (rb = 0: none), kb = 3 leg + group.
template <int LPL>
QPB_HD void direction(const FastParams& K, Lane<LPL>& ln, int j, const double* side, StepTmp<LPL>& T, double& ub, double& rb, int& kb) {
  const double* Gs = side + kSideG;
  if (!ln.done && ln.iters >= K.max_iter) {
    ln.status = QPB_MAX_ITER;
    ln.done = true;
  }
  T.act = !ln.done;
  if (T.act) ln.iters++;
  const int p = T.act ? ln.p : 0;
  const int pl = (p * 11) >> 5, pg = p - 3 * pl;  // leg and group of row p
  double n[3], dp;
  row_normal(K, pg, T.act ? ln.pc : 1u, n, dp);
  // tangential part tn of n on the face of leg pl; a = A_pl tn
  double tn[3];
  {
    const double rx = side[kSideR + 3 * pl], ry = side[kSideR + 3 * pl + 1], rz = side[kSideR + 3 * pl + 2];
    const Leg Lp = leg_of(ln.word >> (6 * pl), (ln.stance >> pl) & 1u, K);
    leg_proj(Lp, n, tn);
    A_t(tn, rx, ry, rz, T.a);
  }
  T.zeta0 = n[0] * tn[0] + n[1] * tn[1] + n[2] * tn[2];  // > 0: one row per group is always independent
  // G yh = a / 2
  double G[21];
#pragma unroll
  for (int i = 0; i < 21; i++) G[i] = Gs[i];
  const bool pd = chol6<true>(G);
  double yh[6];
#pragma unroll
  for (int i = 0; i < 6; i++) yh[i] = 0.5 * T.a[i];
  chol6_solve(G, yh);
  // curvature along the step: zeta = n' z = (zeta0 - 2 a' yh) / (2 w)
  double ayh = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) ayh = fma(T.a[i], yh[i], ayh);
  const double hw = 0.5 * K.inv_w;
  T.zeta = hw * fma(-2.0, ayh, T.zeta0);
  if (T.act && (!pd || !(T.zeta > 0.0) || !(T.zeta0 > 0.0))) {  // cannot happen within the face family
    ln.status = QPB_MAX_ITER;
    ln.done = true;
    T.act = false;
  }
  // own legs: z_i = (1/w) (tn / 2 [i = pl] - Pi_i A_i' yh), N r = n [i = pl] - Q z; blocking candidate
  ub = 1.0;
  rb = 0.0;
  kb = -1;
#pragma unroll
  for (int li = 0; li < LPL; li++) {
    const int i = LPL * j + li;
    const Leg L = leg_of(ln.word >> (6 * i), (ln.stance >> i) & 1u, K);
    double g[3], pgv[3], h[3], m[3];
    At_y(yh, ln.r[3 * li], ln.r[3 * li + 1], ln.r[3 * li + 2], g);
    leg_proj(L, g, pgv);
    const bool me = i == pl;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double zk = -K.inv_w * pgv[k];
      T.z[3 * li + k] = me ? fma(hw, tn[k], zk) : zk;
      const double hk = -2.0 * (g[k] - pgv[k]);  // the component of -2 A' yh normal to the face
      h[k] = me ? hk + (n[k] - tn[k]) : hk;
    }
    leg_multipliers(K, L, h, m);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      T.rr[3 * li + k] = m[k];
      const bool cand = hi32(m[k]) > 0 && ln.u[3 * li + k] * rb < ub * m[k];  // m > 0 and u / m < ub / rb
      if (cand) {
        ub = ln.u[3 * li + k];
        rb = m[k];
        kb = 3 * i + k;
      }
    }
  }
  if (!T.act) {
    rb = 0.0;
    kb = -1;
  }
}

// combine two blocking candidates (exchange 3 is a reduction with this operator)
QPB_HD void better_ratio(double& ub, double& rb, int& kb, double ub2, double rb2, int kb2) {
  if (hi32(rb2) > 0 && ub2 * rb < ub * rb2) {
    ub = ub2;
    rb = rb2;
    kb = kb2;
  }
}

// After exchange 3: step, working-set change, rank-one update of G.  store_G: this lane writes G back (one per QP).
template <int LPL>
QPB_HD void advance(const FastParams& K, Lane<LPL>& ln, int j, double* side, const StepTmp<LPL>& T, double ub, double rb, int kb,
                    bool store_G) {
  double* Gs = side + kSideG;
  const bool has1 = kb >= 0;
  const double t1 = has1 ? ub * rcp_fast(rb) : 0.0;
  const double t2 = -ln.sp * rcp_fast(T.zeta);
  const bool full = !has1 || t2 <= t1;
  const double t = T.act ? (full ? t2 : t1) : 0.0;
#pragma unroll
  for (int i = 0; i < 3 * LPL; i++) {
    ln.f[i] = fma(t, T.z[i], ln.f[i]);
    ln.u[i] = fma(-t, T.rr[i], ln.u[i]);
  }
  ln.up += t;
  ln.sp = fma(t, T.zeta, ln.sp);  // the slack of row p grows by t zeta
  // working-set change: row p enters (full step) or the blocking row leaves (partial step); G follows by a rank-one
  // update with v = A t, t = the row's normal projected on the face WITHOUT that row:  G -+= v v' / (n' t)
  const int p = ln.p < 0 ? 0 : ln.p;
  const int idx = full ? p : (has1 ? kb : 0);
  double v[6], cc;
#pragma unroll
  for (int i = 0; i < 6; i++) v[i] = T.a[i];
  cc = -rcp_fast(T.zeta0);
  uint32_t word = ln.word;
  if (full) {
    word |= ln.pc << (2 * p);
  } else {
    const int kl = (idx * 11) >> 5, kg = idx - 3 * kl;
    const uint32_t kc = (word >> (2 * idx)) & 3u;
    word &= ~(3u << (2 * idx));
    double nk[3], dk, tk[3];
    row_normal(K, kg, kc, nk, dk);
    const double rx = side[kSideR + 3 * kl], ry = side[kSideR + 3 * kl + 1], rz = side[kSideR + 3 * kl + 2];
    const Leg Lk = leg_of(word >> (6 * kl), true, K);
    leg_proj(Lk, nk, tk);
    A_t(tk, rx, ry, rz, v);
    cc = rcp_fast(nk[0] * tk[0] + nk[1] * tk[1] + nk[2] * tk[2]);
  }
  // every lane of the QP updates its own copy of G from the old one; the first lane writes it back once all have read it
  double G[21];
#pragma unroll
  for (int i = 0; i < 21; i++) G[i] = Gs[i];
  rank1(G, v, cc);
  QPB_SYNCWARP();
  if (T.act) {
    ln.word = word;
    const double uv = full ? ln.up : 0.0;
#pragma unroll
    for (int i = 0; i < 3 * LPL; i++)
      if (3 * LPL * j + i == idx) ln.u[i] = uv;
    if (full) ln.p = -1;
    if (store_G) {
#pragma unroll
      for (int i = 0; i < 21; i++) Gs[i] = G[i];
    }
  }
}

// ---- whole-record wrappers: what one thread does before and after the iteration loop ----------------------------

QPB_HD uint32_t stance_mask(uint32_t cbytes) {
  return ((cbytes & 0xffu) ? 1u : 0u) | ((cbytes & 0xff00u) ? 2u : 0u) | ((cbytes & 0xff0000u) ? 4u : 0u) |
         ((cbytes & 0xff000000u) ? 8u : 0u);
}

// rec: slots 0..47 of the state record (attitudes, twists, feet).  hint: bit 31 set = bits 0..23 hold a working set.
// commit(st, G, key) receives every dual-feasible starting pair found (the last call wins) -- or the bad-input record.
template <class Params, class Commit>
QPB_HD void setup(const Params& P, const FastParams& K, const double* rec, uint32_t cbytes, uint32_t hint, State& st,
                  double (&b6)[6], double (&G)[21], Commit& commit) {
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
  if (fin) {
    start(K, st, b6, hint, (hint >> 31) != 0u, G, commit);
  } else {
    st.word = 0u;
    st.iters = 0;
    st.status = QPB_BAD_INPUT;
    st.done = true;
#pragma unroll
    for (int i = 0; i < 12; i++) st.f[i] = st.u[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 21; i++) G[i] = 0.0;
    commit(st, G, 0u);
  }
}

// World-frame solution -> body-frame GRF (balance_controller.cpp:218-232) and tau = J^T f (kinematics.cpp:162-188,
// 218-231, clamp commander_node.cpp:526).  R: Rwb (9), q: joint angles (12).
// UNROLL = 1: one copy of the three sincos expansions instead of four and the per-leg arrays in local memory -- for the
// one-launch kernel, whose active-set loop shares the instruction cache with this code (unrolled, its cold mid-size
// batches lose 15-20 %) and for the early-finish set-up, which has no registers to spare.  UNROLL = 4: everything in
// registers -- for the finishing pass (+7 % on config 3, +3 % on config 2; profiles/r02_finish_unroll_ab.txt).
template <int UNROLL = 1, class Params>
QPB_HD void finish(const Params& P, const double* R, const double* q, const State& st, double (&grf)[12], double (&tau)[12]) {
  const bool good = st.status == QPB_OK;
  const bool qok = st.status != QPB_BAD_INPUT;
#pragma unroll UNROLL
  for (int i = 0; i < 4; i++) {
    const bool on = good && ((st.stance >> i) & 1u);
    const double f0 = st.f[3 * i], f1 = st.f[3 * i + 1], f2 = st.f[3 * i + 2];
    double fb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) fb[k] = on ? -1.0 * (R[k] * f0 + R[3 + k] * f1 + R[6 + k] * f2) : 0.0;
    double s1, c1, s2, c2, s23, c23;
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

// What the set-up and finishing passes read of qpb_params (344 of its 1 904 bytes: no S, no W -- FastParams carries what the
// solver needs of those).  Kernel arguments travel with every launch; with the whole qpb_params by value (2.2 KB per
// launch, two launches per pipeline stage) the host pipeline stalled with four stages in flight.
struct EdgeParams {
  double mass;
  double Ib[9];
  double kff[6], kp_p[3], kd_p[3], kp_w[3], kd_w[3];
  double link[12];
  double tau_min, tau_max;
  int32_t clamp_tau;
};
inline EdgeParams make_edge_params(const qpb_params& P) {
  EdgeParams E;
  E.mass = P.mass;
  for (int i = 0; i < 9; i++) E.Ib[i] = P.Ib[i];
  for (int i = 0; i < 6; i++) E.kff[i] = P.kff[i];
  for (int i = 0; i < 3; i++) {
    E.kp_p[i] = P.kp_p[i];
    E.kd_p[i] = P.kd_p[i];
    E.kp_w[i] = P.kp_w[i];
    E.kd_w[i] = P.kd_w[i];
  }
  for (int i = 0; i < 12; i++) E.link[i] = P.link[i];
  E.tau_min = P.tau_min;
  E.tau_max = P.tau_max;
  E.clamp_tau = P.clamp_tau;
  return E;
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
