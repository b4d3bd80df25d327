// qpb_mpc.cuh -- 10-step convex-MPC ground-reaction-force QP (BASELINE config 4), one CTA per QP.
//
// The reference has no code for this path (README.md:22-26 covers only the instantaneous QP of
// balance_controller.cpp); the formulation is stated in include/qpb200.h (qpb_mpc_*) and restated on the CPU in
// oracle/mpc_oracle.c.  What is kept from the reference: the friction-pyramid rows (balance_controller.cpp:278-289)
// and their bounds (:296-301 stance, :312-316 swing) on every foot and step.
//
// One CTA solves one QP entirely in shared memory (215 KB):
//   * swing foot-steps (f = 0) are compacted away: n = 3 * (#stance foot-steps) <= 120 variables;
//   * the condensed Hessian is never formed through the 130x120 prediction matrix: entry (a, b) is a closed form
//     in the per-variable torque arm g_a = I_k^-1 (r x e_c) and prefix sums of cos/sin(psi_k);
//   * Cholesky and the inverse factor X = L^-1 come out of ONE Gauss-Jordan sweep (n steps, one barrier each,
//     every step a rank-1 update of all columns to the right of the pivot), X^T stored above the diagonal;
//   * the QP is solved by the same whitened operator-form Goldfarb-Idnani dual active-set method as the balance
//     kernel: N* = (N~^T N~)^-1 N~^T kept explicitly, the projector applied as
//     z~ = n~ - X (N r) through the sparse rows, so an iteration is three triangular mat-vecs and a rank-1 update.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qpb200.h"

namespace qpbmpc {

constexpr int NWMAX = 16;  // the kernel is instantiated for 256 and 512 threads per CTA (template parameter NT)
constexpr int NH = 10;
constexpr int NV = 120;   // variables before compaction
constexpr int LD = 121;   // odd leading dimension: row- and column-wise sweeps are both bank-conflict free
constexpr int QMAX = 90;  // rows of N* with their own storage; rows 90..119 live in the dead lower triangle of M
constexpr int QCAP = 120; // working-set capacity = the largest possible number of independent rows
constexpr int QSPARSE = 32; // up to this many active rows z~ = n~ - sum_k r_k n~_k is formed slot by slot (no mat-vec)
constexpr int MROWS = 240;
constexpr unsigned FULL = 0xffffffffu;

struct DevParams {
  double mu, mass, fzmin, fzmax;
  double Ibinv[9];
  double dt;
  double Lw[12];
  double sLw[3];  // sqrt of the three attitude weights
  double alpha;
  int max_iter;
  int pad;
};

struct __align__(16) Smem {
  double M[NV * LD];    // lower: H, later rows 90..119 of N*; strict upper: X^T (X = L^-1)
  double NS[QMAX * NV]; // rows 0..89 of N*; during assembly: attitude table [n][10][3]
  double rec[272];      // staged record; reused as the output record
  double G[NV * 3];     // g_a = I_k^-1 (r x e_c)
  double dinv[NV];      // 1 / L_jj  (= X_jj)
  double f[NV];         // primal iterate = f + f2: the two halves of the step mat-vec accumulate into their own array
  double f2[NV];
  double nt[NV];
  double zt[NV];
  double w[NV];         // gradient during assembly, N r in the loop
  double E[NH * 12];    // weighted free-response error per step
  double Iinv[NH * 9];
  double Cs[NH], Ss[NH];
  double col[2][160];   // Gauss-Jordan sweep: the pivot column, double-buffered; row R at (R & 15) * 10 + (R >> 4)
  double dval[128];     // pivots d_j
  double u[QCAP];
  double r[QCAP];
  double dd[QCAP];
  double ratio[QCAP];   // max(u_k, 0) / r_k where r_k > 0, else +inf
  double scv[QCAP], scz[QCAP];          // per working-set slot: coefficients of its row on the lateral and the z variable
  unsigned char sv[QCAP], sz[QCAP];     // ... and the indices of those two variables
  double part[2][NWMAX];   // per-warp partial sums of |n~|^2 and n~.z~
  int A[QCAP];
  unsigned red[NWMAX];
  unsigned char slot_of_row[MROWS];
  unsigned char act[MROWS];
  unsigned char sfk[40], sff[40];
  int ns;
  unsigned ticket;
};

#ifdef QPB_MPC_PROFILE
// developer build only (tools/time_mpc.py): cycles per phase summed over all QPs, [assembly, sweep, start, loop, io, count]
__device__ unsigned long long g_mpc_prof[24];
#define MPC_TICK(slot)                                              \
  do {                                                              \
    if (tid == 0) {                                                 \
      const long long _now = clock64();                             \
      atomicAdd(&g_mpc_prof[slot], (unsigned long long)(_now - _t)); \
      _t = _now;                                                    \
    }                                                               \
  } while (0)
#else
#define MPC_TICK(slot) \
  do {                 \
  } while (0)
#endif

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// MUFU seed (>= 20 good bits) + one third-order step: relative error <= 2^-60
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

// element i of row k of N*: rows beyond QMAX are stored as two halves of 60 in the lower triangle of M (rows 119..60,
// columns 0..59), which is dead once the sweep has moved H into registers
__device__ __forceinline__ double& ns_at(Smem& S, int k, int i) {
  if (k < QMAX) return S.NS[k * NV + i];
  const int h = i >= 60 ? 1 : 0;
  return S.M[(119 - 2 * (k - QMAX) - h) * LD + (i - 60 * h)];
}

// Register tiling of the sweep: a 16 x 16 thread grid; thread (ty, tx) owns the elements (ty + 16a, tx + 16b), a, b < 8.
// The pivot column travels through shared memory in a permuted layout -- row R at (R & 15) * 10 + (R >> 4) -- so that
// the eight rows a thread owns (and the eight column multipliers it needs) are contiguous: publishing and reading a
// column are 128-bit accesses, and the stride of 10 doubles keeps the quarter-warps of those accesses conflict-free.
constexpr int CPAD = 10;

template <int B, int NB>
__device__ __forceinline__ void publish_col(const double (&m)[8][8], double* dst, int ty) {
  constexpr int BB = B < 8 ? B : 7;
  double2* d2 = reinterpret_cast<double2*>(dst + CPAD * ty);
#pragma unroll
  for (int a = 0; a < NB; a += 2) d2[a >> 1] = make_double2(m[a][BB], m[a + 1][BB]);
}

// Steps j = 16*BJ .. 16*BJ+15 of the Gauss-Jordan sweep over NB row blocks of 16 (see phase E).  The block index of
// the pivot is a template parameter, so every register index is static and eliminated column blocks cost nothing.
template <int NB, int BJ>
struct SweepBlock {
  // one step; CUR = parity of the buffer that holds pivot column j
  template <int CUR>
  static __device__ __forceinline__ void step(double (&m)[8][8], Smem& S, int tj, int tx, int ty, int tid, bool& ok) {
    const double* cur = S.col[CUR];
    double* nxt = S.col[CUR ^ 1];
    const double d = cur[CPAD * tj + BJ];  // M[j][j], j = 16 BJ + tj
    ok = ok && (d > 0.0);
    const double rd = rcp_fast(d);
    if (tid == 0) S.dval[16 * BJ + tj] = d;
    double ci[8], cv[8];
    const double2* cv2 = reinterpret_cast<const double2*>(cur + CPAD * ty);
    const double2* ci2 = reinterpret_cast<const double2*>(cur + CPAD * tx);
#pragma unroll
    for (int a = 0; a < NB; a += 2) {
      const double2 v = cv2[a >> 1];
      cv[a] = v.x;
      cv[a + 1] = v.y;
    }
#pragma unroll
    for (int b = BJ & ~1; b < NB; b += 2) {  // rows/columns >= n are zero already
      const double2 v = ci2[b >> 1];
      ci[b] = v.x * rd;
      ci[b + 1] = v.y * rd;
    }
    const bool rowj = ty == tj;
    if (rowj) cv[BJ] = 0.0;         // row j itself is assigned below
    if (tx <= tj) ci[BJ] = 0.0;     // columns <= j are finished
    // Row blocks a <= BJ hold rows of the inverse (all columns right of the pivot change); row blocks a > BJ hold the
    // trailing Hessian, of which only the lower triangle (b <= a) is ever read again -- the tiles above it are skipped
    // and simply overwritten when their rows become pivot rows.
#pragma unroll
    for (int b = BJ; b < NB; b++)
#pragma unroll
      for (int a = 0; a < NB; a++)
        if (a <= BJ || b <= a) m[a][b] = fma(-cv[a], ci[b], m[a][b]);
    if (rowj) {
      if (tx > tj) m[BJ][BJ] = -ci[BJ];
#pragma unroll
      for (int b = BJ + 1; b < NB; b++) m[BJ][b] = -ci[b];
    }
    if (tj < 15) {
      if (tx == tj + 1) publish_col<BJ, NB>(m, nxt, ty);
    } else if (BJ + 1 < NB) {
      if (tx == 0) publish_col<BJ + 1, NB>(m, nxt, ty);
    }
    __syncthreads();
  }
  static __device__ __forceinline__ bool run(double (&m)[8][8], Smem& S, int n, int tx, int ty, int tid, bool ok) {
#pragma unroll 1
    for (int tj = 0; tj < 16; tj += 2) {  // 16 BJ is even: even steps read buffer 0, odd steps buffer 1
      if (16 * BJ + tj >= n) return ok;
      step<0>(m, S, tj, tx, ty, tid, ok);
      if (16 * BJ + tj + 1 >= n) return ok;
      step<1>(m, S, tj + 1, tx, ty, tid, ok);
    }
    return SweepBlock<NB, BJ + 1>::run(m, S, n, tx, ty, tid, ok);
  }
};
template <int NB>
struct SweepBlock<NB, NB> {
  static __device__ __forceinline__ bool run(double (&)[8][8], Smem&, int, int, int, int, bool ok) { return ok; }
};

// load the symmetric H into the register tiles, sweep, and store X^T above the diagonal of S.M
template <int NB>
__device__ __noinline__ bool sweep(Smem& S, int n, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
  double m[8][8];
#pragma unroll
  for (int a = 0; a < NB; a++)
#pragma unroll
    for (int b = 0; b < NB; b++) {
      const int R = ty + 16 * a, C = tx + 16 * b;
      m[a][b] = (R < n && C < n) ? (C <= R ? S.M[R * LD + C] : S.M[C * LD + R]) : 0.0;
    }
  if (tx == 0) publish_col<0, NB>(m, S.col[0], ty);
  __syncthreads();
  // a non-positive pivot (H not positive definite: only possible with non-finite or absurd inputs) poisons the rest of
  // the sweep with NaN/Inf but cannot hang it; it is reported once at the end
  const bool ok = SweepBlock<NB, 0>::run(m, S, n, tx, ty, tid, true);
  if (!__syncthreads_and(ok)) return false;
  if (tid < n) S.dinv[tid] = 1.0 / sqrt(S.dval[tid]);
  __syncthreads();
#pragma unroll
  for (int b = 0; b < NB; b++) {
    const int C = tx + 16 * b;
    const double sc = C < n ? S.dinv[C] : 0.0;
#pragma unroll
    for (int a = 0; a < NB; a++) {
      const int R = ty + 16 * a;
      if (C > R && C < n) S.M[R * LD + C] = m[a][b] * sc;  // X[i][c] = Xu[i][c] / L_ii at M[c][i]
    }
  }
  return true;
}

// X(i, c) for the inverse factor stored transposed above the diagonal
__device__ __forceinline__ double X_at(const Smem& S, int i, int c) {
  return i > c ? S.M[c * LD + i] : (i == c ? S.dinv[i] : 0.0);
}

// Triangular mat-vecs with X^T stored above the diagonal of M (leading dimension 121).  Mapping: a warp serves the
// pair of RB-row blocks (warp, nblk-1-warp) -- a short and a long one, n-1 terms together, so the warps are balanced --
// with RB = 8 rows per block for 8 warps and 4 for 16 warps.  lane = RB*s + rr: rr picks the row of the block, s the
// residue class of the inner index; each lane takes four consecutive inner indices per window of W = 4 * (32/RB), and
// within a half-warp the classes sit 8 (RB = 8) or 4 (RB = 4) inner indices apart, which with the odd leading dimension
// makes every 64-bit shared-memory access conflict-free in both sweep directions.  Partial sums are combined over the
// classes with shuffles; every lane returns the full sum.
template <int NT>
struct Tri {
  static constexpr int RB = NT == 256 ? 8 : 4;  // rows per block
  static constexpr int NCL = 32 / RB;           // residue classes
  static constexpr int W = 4 * NCL;             // inner window
};
struct TriMap {
  int iA, iB, off;
  bool onA, onB;
};
template <int NT>
__device__ __forceinline__ TriMap tri_map(int n, int lane, int warp) {
  constexpr int RB = Tri<NT>::RB;
  const int nblk = (n + RB - 1) / RB, bA = warp, bB = nblk - 1 - warp, rr = lane % RB, s = lane / RB;
  TriMap t;
  t.iA = RB * bA + rr;
  t.iB = RB * bB + rr;
  t.off = RB == 8 ? 8 * (s & 1) + 4 * (s >> 1) : 4 * (s & 3) + 16 * (s >> 2);
  t.onA = bA <= bB && t.iA < n;
  t.onB = bA < bB && t.iB < n;
  return t;
}
template <int NT>
__device__ __forceinline__ double class_sum(double a) {
  if (Tri<NT>::RB == 4) a += __shfl_xor_sync(FULL, a, 4);
  a += __shfl_xor_sync(FULL, a, 8);
  return a + __shfl_xor_sync(FULL, a, 16);
}
// sum_{c < i} M[c][i] v_c over this lane's residue class
template <int NT>
__device__ __forceinline__ double tri_col_part(const double* Mi, const double* v, int i, int off) {
  constexpr int W = Tri<NT>::W;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int c = off;
  for (; c + 3 < i; c += W) {
    a0 = fma(Mi[c * LD], v[c], a0);
    a1 = fma(Mi[(c + 1) * LD], v[c + 1], a1);
    a2 = fma(Mi[(c + 2) * LD], v[c + 2], a2);
    a3 = fma(Mi[(c + 3) * LD], v[c + 3], a3);
  }
  if (c < i) a0 = fma(Mi[c * LD], v[c], a0);
  if (c + 1 < i) a1 = fma(Mi[(c + 1) * LD], v[c + 1], a1);
  if (c + 2 < i) a2 = fma(Mi[(c + 2) * LD], v[c + 2], a2);
  return (a0 + a1) + (a2 + a3);
}
// sum_{c < i < n} M[c][i] v_i over this lane's residue class
template <int NT>
__device__ __forceinline__ double tri_row_part(const double* Mc, const double* v, int c, int n, int off) {
  constexpr int W = Tri<NT>::W;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = ((c + 1) & ~(W - 1)) + off;  // first W-aligned window that can hold an index > c
  if (i + 3 > c) {                     // head window: some of its four indices may still be <= c
    if (i > c && i < n) a0 = fma(Mc[i], v[i], a0);
    if (i + 1 > c && i + 1 < n) a1 = fma(Mc[i + 1], v[i + 1], a1);
    if (i + 2 > c && i + 2 < n) a2 = fma(Mc[i + 2], v[i + 2], a2);
    if (i + 3 < n) a3 = fma(Mc[i + 3], v[i + 3], a3);
  }
  i += W;
  for (; i + 3 < n; i += W) {
    a0 = fma(Mc[i], v[i], a0);
    a1 = fma(Mc[i + 1], v[i + 1], a1);
    a2 = fma(Mc[i + 2], v[i + 2], a2);
    a3 = fma(Mc[i + 3], v[i + 3], a3);
  }
  if (i < n) a0 = fma(Mc[i], v[i], a0);
  if (i + 1 < n) a1 = fma(Mc[i + 1], v[i + 1], a1);
  if (i + 2 < n) a2 = fma(Mc[i + 2], v[i + 2], a2);
  return (a0 + a1) + (a2 + a3);
}
// (X v)_i = sum_{c < i} M[c][i] v_c + dinv_i v_i for rows iA, iB of the map
template <int NT>
__device__ __forceinline__ void tri_X8(const Smem& S, const double* v, const TriMap& t, double& oA, double& oB) {
  const double a = class_sum<NT>(t.onA ? tri_col_part<NT>(&S.M[t.iA], v, t.iA, t.off) : 0.0);
  const double b = class_sum<NT>(t.onB ? tri_col_part<NT>(&S.M[t.iB], v, t.iB, t.off) : 0.0);
  oA = t.onA ? fma(S.dinv[t.iA], v[t.iA], a) : 0.0;
  oB = t.onB ? fma(S.dinv[t.iB], v[t.iB], b) : 0.0;
}
// (X^T v)_c = dinv_c v_c + sum_{i > c} M[c][i] v_i for columns iA, iB of the map
template <int NT>
__device__ __forceinline__ void tri_XT8(const Smem& S, const double* v, int n, const TriMap& t, double& oA, double& oB) {
  const double a = class_sum<NT>(t.onA ? tri_row_part<NT>(&S.M[t.iA * LD], v, t.iA, n, t.off) : 0.0);
  const double b = class_sum<NT>(t.onB ? tri_row_part<NT>(&S.M[t.iB * LD], v, t.iB, n, t.off) : 0.0);
  oA = t.onA ? fma(S.dinv[t.iA], v[t.iA], a) : 0.0;
  oB = t.onB ? fma(S.dinv[t.iB], v[t.iB], b) : 0.0;
}

// sum over the 16 lanes of a half-warp
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// exact minimum over the warp of non-negative doubles (or +inf): their bit patterns order like unsigned integers
__device__ __forceinline__ double warp_min_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v);
  const unsigned mh = __reduce_min_sync(FULL, hi);
  const unsigned lo = (hi == mh) ? (unsigned)__double2loint(v) : 0xffffffffu;
  const unsigned ml = __reduce_min_sync(FULL, lo);
  return __hiloint2double((int)mh, (int)ml);
}

// coefficient of pyramid row type t (0..5) on component comp (0..2) of its foot-step
__device__ __forceinline__ double row_coef(int t, int comp, double mu) {
  if (comp == 2) return t < 4 ? mu : (t == 4 ? 1.0 : -1.0);
  if (comp == 0) return t == 0 ? -1.0 : (t == 3 ? 1.0 : 0.0);
  return t == 1 ? -1.0 : (t == 2 ? 1.0 : 0.0);
}

// slack n.f - b of row type t, branch-free
__device__ __forceinline__ double row_slack(int t, double fx, double fy, double fz, const DevParams& P) {
  const double lat = (t == 0 || t == 3) ? fx : fy;                    // lateral component the row looks at
  const double sl = (t == 0 || t == 1) ? -lat : lat;
  const double pyr = fma(P.mu, fz, sl);
  return t < 4 ? pyr : (t == 4 ? fz - P.fzmin : P.fzmax - fz);
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
mpc_qp_kernel(const __grid_constant__ DevParams P, const qpb_mpc_rec* __restrict__ in, qpb_mpc_out_rec* __restrict__ out,
              int64_t nrec, unsigned long long* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  int64_t rec = blockIdx.x;
  while (rec < nrec) {
    // ---- stage the record; draw the next ticket early -------------------------------------------------------
#ifdef QPB_MPC_PROFILE
    long long _t = clock64();
#endif
    if (tid == 0) S.ticket = (unsigned)atomicAdd(ticket, 1ull);
    if (tid < 136) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(in + rec) + tid);
      S.rec[2 * tid] = v.x;
      S.rec[2 * tid + 1] = v.y;
    }
    for (int i = tid; i < MROWS; i += NT) S.act[i] = 0;
    __syncthreads();
    int bad = 0;
    for (int i = tid; i < 263; i += NT) bad |= !isfinite(S.rec[i]);
    bad = __syncthreads_or(bad);
    int status = bad ? QPB_BAD_INPUT : QPB_OK, iters = 0, n = 0;

    if (!bad) {
      MPC_TICK(16);
      // ---- phase A: compaction of the stance foot-steps (warp 0); per-step trigonometry and inertia (warp 1) --
      const unsigned char* cb = reinterpret_cast<const unsigned char*>(&S.rec[263]);
      if (warp == 0) {
        const bool c1 = cb[lane] != 0, c2 = lane < 8 && cb[32 + lane] != 0;
        const unsigned m1 = __ballot_sync(FULL, c1), m2 = __ballot_sync(FULL, c2);
        const int n1 = __popc(m1);
        const unsigned below = (1u << lane) - 1u;
        if (c1) {
          const int pos = __popc(m1 & below);
          S.sfk[pos] = (unsigned char)(lane >> 2);
          S.sff[pos] = (unsigned char)(lane & 3);
        }
        if (c2) {
          const int pos = n1 + __popc(m2 & below);
          S.sfk[pos] = (unsigned char)((32 + lane) >> 2);
          S.sff[pos] = (unsigned char)(lane & 3);
        }
        if (lane == 0) S.ns = n1 + __popc(m2);
      } else if (warp == 1) {
        double c = 0.0, s = 0.0;
        if (lane < NH) sincos(S.rec[13 + 13 * lane + 2], &s, &c);
        double pc = c, ps = s;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const double tc = __shfl_up_sync(FULL, pc, o), ts = __shfl_up_sync(FULL, ps, o);
          if (lane >= o) { pc += tc; ps += ts; }
        }
        if (lane < NH) {
          S.Cs[lane] = pc;
          S.Ss[lane] = ps;
          // I_k^-1 = Rz Ib^-1 Rz^T
          const double Rz[9] = { c, -s, 0.0, s, c, 0.0, 0.0, 0.0, 1.0 };
          double t[9];
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
              t[3 * i + j] = Rz[3 * i] * P.Ibinv[j] + Rz[3 * i + 1] * P.Ibinv[3 + j] + Rz[3 * i + 2] * P.Ibinv[6 + j];
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
              S.Iinv[9 * lane + 3 * i + j] = t[3 * i] * Rz[3 * j] + t[3 * i + 1] * Rz[3 * j + 1] + t[3 * i + 2] * Rz[3 * j + 2];
        }
      }
      __syncthreads();
      const int ns = S.ns;
      n = 3 * ns;
      const double dt = P.dt, dt2 = dt * dt, im = 1.0 / P.mass;

      MPC_TICK(17);
      // ---- phase B: torque arms g_a (threads < n); weighted free-response error (threads 128..137) ----------
      if (tid < n) {
        const int c = tid / 3, comp = tid - 3 * c, k = S.sfk[c], foot = S.sff[c];
        const double* r = &S.rec[143 + 12 * k + 3 * foot];
        const double rx0 = comp == 1 ? -r[2] : (comp == 2 ? r[1] : 0.0);
        const double rx1 = comp == 0 ? r[2] : (comp == 2 ? -r[0] : 0.0);
        const double rx2 = comp == 0 ? -r[1] : (comp == 1 ? r[0] : 0.0);
        const double* I = &S.Iinv[9 * k];
#pragma unroll
        for (int i = 0; i < 3; i++) S.G[3 * tid + i] = I[3 * i] * rx0 + I[3 * i + 1] * rx1 + I[3 * i + 2] * rx2;
      } else if (tid >= 128 && tid < 128 + NH) {
        const int k = tid - 128;
        const double* x0 = S.rec;
        const double* xr = &S.rec[13 + 13 * k];
        const double C = S.Cs[k], Sn = S.Ss[k], k1 = (double)(k + 1);
        const double w0 = x0[6], w1 = x0[7], w2 = x0[8], g = x0[12];
        double fr[12];
        fr[0] = x0[0] + dt * (C * w0 + Sn * w1);
        fr[1] = x0[1] + dt * (-Sn * w0 + C * w1);
        fr[2] = x0[2] + dt * k1 * w2;
        fr[3] = x0[3] + k1 * dt * x0[9];
        fr[4] = x0[4] + k1 * dt * x0[10];
        fr[5] = x0[5] + k1 * dt * x0[11] + dt2 * g * (0.5 * k * (k + 1));
        fr[6] = w0; fr[7] = w1; fr[8] = w2;
        fr[9] = x0[9]; fr[10] = x0[10];
        fr[11] = x0[11] + k1 * dt * g;
#pragma unroll
        for (int s = 0; s < 12; s++) S.E[12 * k + s] = P.Lw[s] * (fr[s] - xr[s]);
      }
      __syncthreads();

      MPC_TICK(18);
      // ---- phase C: per-step-pair kernels of the Hessian and per-step vectors of the gradient -------------------
      // With P_k = sum_{l<=k} T_l and S_kj = P_k - P_j (the attitude response of step k to a torque impulse at step j),
      //   H[a][b] = 2 g_a' Kf(ja, jb) g_b + (a diagonal term when a and b are the same force component) + 2 alpha [a==b],
      //   Kf(ja, jb) = dt^4 sum_{k>ja} S_kja' diag(Lw_att) S_kjb + dt^2 (10 - ja) diag(Lw_omega)        (jb <= ja)
      //   gvec[a] = 2 ( g_a' (qa(ja) + dt qw(ja)) + dt^2/m qp(ja)[c] + dt/m qv(ja)[c] ),
      //   qa(j) = dt^2 sum_{k>j} S_kj' e_att(k), qw(j) = sum_{k>=j} e_omega(k), qp(j) = sum_{k>=j} (k-j) e_p(k), qv(j) = sum_{k>=j} e_v(k)
      // with e(k) the weighted free-response error of step k.  55 step pairs and 10 steps: no per-variable table at all.
      double* Kf = S.NS;             // [ja][jb][9]
      double* Qv = S.NS + 900;       // [j][12]: qa + dt qw | qp | qv | (unused)
      if (tid < 100) {
        const int ja = tid / NH, jb = tid - NH * ja;
        if (jb <= ja) {
          double k00 = 0.0, k01 = 0.0, k10 = 0.0, k11 = 0.0, k22 = 0.0;
          const double Ca0 = S.Cs[ja], Sa0 = S.Ss[ja], Cb0 = S.Cs[jb], Sb0 = S.Ss[jb];
          for (int k = ja + 1; k < NH; k++) {
            const double Ck = S.Cs[k], Sk = S.Ss[k];
            const double Ca = Ck - Ca0, Sa = Sk - Sa0, Cb = Ck - Cb0, Sb = Sk - Sb0;
            k00 += P.Lw[0] * Ca * Cb + P.Lw[1] * Sa * Sb;
            k01 += P.Lw[0] * Ca * Sb - P.Lw[1] * Sa * Cb;
            k10 += P.Lw[0] * Sa * Cb - P.Lw[1] * Ca * Sb;
            k11 += P.Lw[0] * Sa * Sb + P.Lw[1] * Ca * Cb;
            k22 += P.Lw[2] * (double)((k - ja) * (k - jb));
          }
          const double dt4 = dt2 * dt2, cw = dt2 * (double)(NH - ja);
          double* o = &Kf[9 * tid];
          o[0] = dt4 * k00 + cw * P.Lw[6]; o[1] = dt4 * k01;               o[2] = 0.0;
          o[3] = dt4 * k10;               o[4] = dt4 * k11 + cw * P.Lw[7]; o[5] = 0.0;
          o[6] = 0.0;                     o[7] = 0.0;                     o[8] = dt4 * k22 + cw * P.Lw[8];
        }
      } else if (tid >= 128 && tid < 128 + NH) {
        const int j = tid - 128;
        double qa0 = 0.0, qa1 = 0.0, qa2 = 0.0, qw[3] = { 0.0, 0.0, 0.0 }, qp[3] = { 0.0, 0.0, 0.0 }, qv[3] = { 0.0, 0.0, 0.0 };
        for (int k = j; k < NH; k++) {
          const double* e = &S.E[12 * k];
          const double C = S.Cs[k] - S.Cs[j], Sn = S.Ss[k] - S.Ss[j], dk = (double)(k - j);
          qa0 += C * e[0] - Sn * e[1];  // S_kj' e_att  (zero for k == j)
          qa1 += Sn * e[0] + C * e[1];
          qa2 += dk * e[2];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            qp[c] += dk * e[3 + c];
            qw[c] += e[6 + c];
            qv[c] += e[9 + c];
          }
        }
        double* o = &Qv[12 * j];
        o[0] = dt2 * qa0 + dt * qw[0];
        o[1] = dt2 * qa1 + dt * qw[1];
        o[2] = dt2 * qa2 + dt * qw[2];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          o[3 + c] = dt2 * im * qp[c];
          o[6 + c] = dt * im * qv[c];
        }
      }
      __syncthreads();

      MPC_TICK(19);
      // ---- phase D: H in 3x3 blocks (foot-step fa x foot-step fb, fb <= fa) and the gradient ------------------
      // Rows fa' and ns-1-fa' of the block triangle together hold ns+1 blocks, so idx -> (fa, fb) needs one division
      // and no thread draws a block above the diagonal.
      {
        const int per = ns + 1, half = (ns + 1) >> 1;
        for (int idx = tid; idx < half * per; idx += NT) {
          const int rp = idx / per, t = idx - rp * per;
          const bool lowrow = t <= rp;
          if (!lowrow && 2 * rp == ns - 1) continue;  // odd ns: the middle row is its own partner
          const int fa = lowrow ? rp : ns - 1 - rp, fb = lowrow ? t : t - rp - 1;
          const int ja = S.sfk[fa], jb = S.sfk[fb];  // jb <= ja: the compact order is step-major
          const double* K = &Kf[9 * (NH * ja + jb)];
          const double k00 = K[0], k01 = K[1], k10 = K[3], k11 = K[4], k22 = K[8];
          double A[3][3], T[3][3];  // T[l] = Kf g_b,l
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const double b0 = S.G[9 * fb + 3 * i], b1 = S.G[9 * fb + 3 * i + 1], b2 = S.G[9 * fb + 3 * i + 2];
            T[i][0] = k00 * b0 + k01 * b1;
            T[i][1] = k10 * b0 + k11 * b1;
            T[i][2] = k22 * b2;
#pragma unroll
            for (int c = 0; c < 3; c++) A[i][c] = S.G[9 * fa + 3 * i + c];
          }
          const double cnt = (double)(NH - ja);
          const double s1 = 0.5 * (cnt - 1.0) * cnt, s2 = (cnt - 1.0) * cnt * (2.0 * cnt - 1.0) * (1.0 / 6.0);
          const double lin = (dt * im) * (dt * im) * cnt, quad = (dt2 * im) * (dt2 * im) * (s2 + (double)(ja - jb) * s1);
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int l = 0; l < 3; l++) {
              double v = A[i][0] * T[l][0] + A[i][1] * T[l][1] + A[i][2] * T[l][2];
              if (i == l) {
                v += P.Lw[9 + i] * lin + P.Lw[3 + i] * quad;
                if (fa == fb) v += P.alpha;
              }
              S.M[(3 * fa + i) * LD + 3 * fb + l] = 2.0 * v;  // diagonal blocks: all 9
            }
        }
      }
      MPC_TICK(20);
      if (tid < n) {
        const int a = tid, fa = a / 3, ca = a - 3 * fa, ja = S.sfk[fa];
        const double* ga = &S.G[3 * a];
        const double* qj = &Qv[12 * ja];
        S.w[a] = 2.0 * (ga[0] * qj[0] + ga[1] * qj[1] + ga[2] * qj[2] + qj[3 + ca] + qj[6 + ca]);
      }
      __syncthreads();

      MPC_TICK(0);
      // ---- phase E: Gauss-Jordan sweep on a register-resident matrix ------------------------------------------
      // Thread (ty, tx) of a 16x16 grid owns the elements (ty + 16a, tx + 16b) of the symmetric H.  Step j reads pivot
      // column j from shared memory (published by its owners at the end of step j-1) and applies
      //   M[R][C] -= M[R][j] * M[C][j] / d_j   to every column C > j of every row R != j,   M[j][C] = -M[C][j] / d_j.
      // Below the diagonal this is the Cholesky elimination (never read again); above it, it accumulates the rows of
      // the unscaled inverse factor: afterwards X[i][c] = M[c][i] / sqrt(d_i) for c < i.  One barrier per step.
      {
        const int nb = (n + 15) >> 4;
        bool ok;
        static_assert(NT == 256, "the sweep is written for a 16 x 16 thread grid");
        if (nb <= 2) ok = sweep<2>(S, n, tid);
        else if (nb <= 4) ok = sweep<4>(S, n, tid);
        else if (nb <= 6) ok = sweep<6>(S, n, tid);
        else ok = sweep<8>(S, n, tid);
        if (!ok) status = QPB_BAD_INPUT;
      }
    }

    MPC_TICK(1);
    if (status == QPB_OK && n > 0) {
      __syncthreads();
      // ---- unconstrained minimiser f0 = -X^T X g ----------------------------------------------------------
      const TriMap tm = tri_map<NT>(n, lane, warp);
      const bool wrA = tm.onA && lane < Tri<NT>::RB, wrB = tm.onB && lane < Tri<NT>::RB;  // class-0 lanes write the results
      {
        double ya, yb;
        tri_X8<NT>(S, S.w, tm, ya, yb);
        if (wrA) S.zt[tm.iA] = -ya;
        if (wrB) S.zt[tm.iB] = -yb;
        __syncthreads();
        tri_XT8<NT>(S, S.zt, n, tm, ya, yb);
        if (wrA) S.f[tm.iA] = ya;
        if (wrB) S.f[tm.iB] = yb;
        if (tid < NV) S.f2[tid] = 0.0;
        __syncthreads();
      }

      MPC_TICK(2);
      // ---- dual active-set loop ---------------------------------------------------------------------------
      const int m = 2 * n;  // 6 rows per stance foot-step
      const double fzs = 1.0 + fmax(fabs(P.fzmin), fabs(P.fzmax));
      const int hl = lane & 15, hh = lane >> 4;         // half-warp per working-set slot
      const int ci = tid & 127, ck = tid >> 7;          // rank-1 updates: column ci, slots ck, ck + NT/128, ...
      constexpr int CKS = NT / 128;
      int q = 0;
      for (;;) {
        unsigned key = 0;
        if (tid < m && S.act[tid] == 0) {
          const int c = tid / 6, t = tid - 6 * c;
          const double s = row_slack(t, S.f[3 * c] + S.f2[3 * c], S.f[3 * c + 1] + S.f2[3 * c + 1], S.f[3 * c + 2] + S.f2[3 * c + 2], P);
          const double tol = t < 4 ? 1e-9 : 1e-9 * fzs;
          if (s < -tol) key = ((unsigned)__double2hiint(-s) & 0xffffff00u) | (unsigned)tid;
        }
        key = __reduce_max_sync(FULL, key);
        if (lane == 0) S.red[warp] = key;
        __syncthreads();
        key = S.red[0];
#pragma unroll
        for (int i = 1; i < NW; i++) key = max(key, S.red[i]);
        MPC_TICK(6);
        if (key == 0u) break;  // primal feasible: optimal
        const int p = (int)(key & 0xffu);
        const int pc = p / 6, pt = p - 6 * pc;
        const int zz = 3 * pc + 2, vv = (pt == 0 || pt == 3) ? 3 * pc : ((pt == 1 || pt == 2) ? 3 * pc + 1 : zz);
        const double cvv = pt < 4 ? row_coef(pt, vv - 3 * pc, P.mu) : 0.0, czz = row_coef(pt, 2, P.mu);
        double sp = row_slack(pt, S.f[3 * pc] + S.f2[3 * pc], S.f[3 * pc + 1] + S.f2[3 * pc + 1], S.f[3 * pc + 2] + S.f2[3 * pc + 2], P);
        double up = 0.0;
        bool stop = false;
        for (;;) {  // steps towards row p until it joins the working set
          if (iters >= P.max_iter) { status = QPB_MAX_ITER; stop = true; break; }
          iters++;
          // (1) n~ = X n_p, per-warp partials of |n~|^2
          {
            double v = 0.0;
            if (tid < n) {
              v = cvv * X_at(S, tid, vv) + czz * X_at(S, tid, zz);
              S.nt[tid] = v;
            }
            if (q == 0) {  // with an empty working set zeta = |n~|^2 is needed right away; otherwise stage (3) forms it
              v = warp_sum(v * v);
              if (lane == 0) S.part[0][warp] = v;
            }
          }
          __syncthreads();
          MPC_TICK(8);
          const double* ztp = S.nt;
          if (q > 0) {
            // (2) r = N* n~ (half-warp per slot); the slot leader also forms the ratio-test entry
            for (int k0 = 2 * warp; k0 < q; k0 += 2 * NW) {
              const int k = k0 + hh;
              const bool valid = k < q;
              double a = 0.0;
              if (valid) {
                if (k < QMAX) {  // the row has its own storage: eight masked terms per lane, loads issued up front
                  const double* row = &S.NS[k * NV];
                  double a1 = 0.0;
#pragma unroll
                  for (int j = 0; j < 8; j += 2) {
                    const int i0 = hl + 16 * j, i1 = i0 + 16;
                    const double m0 = i0 < n ? row[i0] : 0.0, m1 = i1 < n ? row[i1] : 0.0;
                    const double v0 = i0 < n ? S.nt[i0] : 0.0, v1 = i1 < n ? S.nt[i1] : 0.0;
                    a = fma(m0, v0, a);
                    a1 = fma(m1, v1, a1);
                  }
                  a += a1;
                } else {
                  for (int i = hl; i < n; i += 16) a = fma(ns_at(S, k, i), S.nt[i], a);
                }
              }
              a = half_sum(a);
              if (valid && hl == 0) {
                S.r[k] = a;
                S.ratio[k] = a > 0.0 ? fmax(S.u[k], 0.0) * rcp_fast(a) : INF;
              }
            }
            __syncthreads();
            MPC_TICK(9);
            double zp = 0.0, np = 0.0;
            if (q <= QSPARSE) {
              // (3a) few active rows: z~_i = n~_i - sum_k r_k n~_k[i] with n~_k = X n_k rebuilt from two columns of X;
              //      two threads per i (slot parity), combined in a fixed order
              const int i = tid >> 1, par = tid & 1;
              double acc = 0.0;
              if (i < n)
                for (int k = par; k < q; k += 2)
                  acc = fma(S.r[k], fma(S.scv[k], X_at(S, i, S.sv[k]), S.scz[k] * X_at(S, i, S.sz[k])), acc);
              acc += __shfl_xor_sync(FULL, acc, 1);
              if (i < n && par == 0) {
                const double nti = S.nt[i], z = nti - acc;
                S.zt[i] = z;
                zp = z * nti;
                np = nti * nti;
              }
            } else {
              // (3b) w = N r gathered through the sparse rows in a fixed order (an atomic scatter would be one barrier
              //      cheaper but not bitwise reproducible), then z~ = n~ - X w
              if (tid < n) {
                const int c = tid / 3, comp = tid - 3 * c;
                double a = 0.0;
#pragma unroll
                for (int t = 0; t < 6; t++) {
                  const int row = 6 * c + t;
                  if (S.act[row] == 1) a = fma(row_coef(t, comp, P.mu), S.r[S.slot_of_row[row]], a);
                }
                S.w[tid] = a;
              }
              __syncthreads();
              MPC_TICK(10);
              double xa, xb;
              tri_X8<NT>(S, S.w, tm, xa, xb);
              if (wrA) {
                const double na = S.nt[tm.iA], za = na - xa;
                S.zt[tm.iA] = za;
                zp = za * na;
                np = na * na;
              }
              if (wrB) {
                const double nb_ = S.nt[tm.iB], zb = nb_ - xb;
                S.zt[tm.iB] = zb;
                zp = fma(zb, nb_, zp);
                np = fma(nb_, nb_, np);
              }
            }
            // per-warp partials of zeta = n~.z~ and of |n~|^2 (two independent shuffle chains)
            zp = warp_sum(zp);
            np = warp_sum(np);
            if (lane == 0) {
              S.part[1][warp] = zp;
              S.part[0][warp] = np;
            }
            __syncthreads();
            ztp = S.zt;
          }
          MPC_TICK(11);
          // (4) zeta, |n~|^2 (fixed summation order: identical in every thread) and the ratio test
          double nn = 0.0, zeta = 0.0;
#pragma unroll
          for (int i = 0; i < NW; i++) nn += S.part[0][i];
          if (q > 0) {
#pragma unroll
            for (int i = 0; i < NW; i++) zeta += S.part[1][i];
          } else {
            zeta = nn;
          }
          const bool dep = !(zeta > 1e-13 * nn);
          double t1 = INF;
          int ks = -1;
          for (int k = lane; k < q; k += 32) {
            const double rt = S.ratio[k];
            if (rt < t1) { t1 = rt; ks = k; }
          }
          {
            const double tmin = warp_min_nonneg(t1);
            const unsigned who = __ballot_sync(FULL, ks >= 0 && t1 == tmin);
            ks = who ? __shfl_sync(FULL, ks, __ffs(who) - 1) : -1;
            t1 = tmin;
          }
          const bool has1 = ks >= 0;
          if (dep && !has1) {
            // row p lies in the span of the working set and no multiplier can give way: a rounding artefact of a twin
            // row (fzmin == fzmax) is set aside; anything larger is reported
            if (sp < -1e-6 * fzs) { status = QPB_BAD_INPUT; stop = true; }
            if (tid == 0) S.act[p] = 2;
            __syncthreads();
            break;
          }
          const double iz = dep ? 0.0 : rcp_fast(zeta);
          const double t2 = dep ? INF : -sp * iz;
          const bool full = !dep && (!has1 || t2 <= t1);
          const double t = full ? t2 : t1;
          MPC_TICK(12);
          // (5) primal and dual step
          if (!dep) {
            // delta f = X^T z~, one lane per column and half of the inner indices (even / odd): no cross-lane reduction,
            // z~_i is a broadcast load, rows of M are read with the conflict-free stride LD.  Column block wc has about
            // (n - 32 wc) / 2 terms per lane, and warps w and w + 4 share a scheduler, so warp w takes block w and warp
            // w + 4 block 3 - w: every scheduler gets a long and a short one.  Each half accumulates into its own copy of
            // the iterate (f and f2).
            {
              const int hf = warp >> 2, wc = hf ? 3 - (warp & 3) : warp, c0 = 32 * wc, c = c0 + lane;
              double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
              if (c0 < n) {
                const double* row = &S.M[(c < n ? c : 0) * LD];
                int i = c0 + 1 + ((c0 + 1 + hf) & 1);  // first inner index > c0 with the parity of this half
                // head: lanes join as i passes their column; entries at or left of the diagonal (finite leftovers of the
                // assembly) are loaded unconditionally and masked to zero so that the loads pipeline
                for (; i <= c0 + 32 && i + 6 < n; i += 8) {
                  const double m0 = row[i], m1 = row[i + 2], m2 = row[i + 4], m3 = row[i + 6];
                  a0 = fma(i > c ? m0 : 0.0, ztp[i], a0);
                  a1 = fma(i + 2 > c ? m1 : 0.0, ztp[i + 2], a1);
                  a2 = fma(i + 4 > c ? m2 : 0.0, ztp[i + 4], a2);
                  a3 = fma(i + 6 > c ? m3 : 0.0, ztp[i + 6], a3);
                }
                for (; i + 6 < n; i += 8) {
                  a0 = fma(row[i], ztp[i], a0);
                  a1 = fma(row[i + 2], ztp[i + 2], a1);
                  a2 = fma(row[i + 4], ztp[i + 4], a2);
                  a3 = fma(row[i + 6], ztp[i + 6], a3);
                }
                {  // tail: at most three more indices of this parity
                  const double m0 = i < n ? row[i] : 0.0, m1 = i + 2 < n ? row[i + 2] : 0.0, m2 = i + 4 < n ? row[i + 4] : 0.0;
                  const double v0 = i < n ? ztp[i] : 0.0, v1 = i + 2 < n ? ztp[i + 2] : 0.0, v2 = i + 4 < n ? ztp[i + 4] : 0.0;
                  a0 = fma(i > c ? m0 : 0.0, v0, a0);
                  a1 = fma(i + 2 > c ? m1 : 0.0, v1, a1);
                  a2 = fma(i + 4 > c ? m2 : 0.0, v2, a2);
                }
              }
              if (c < n) {
                double a = (a0 + a1) + (a2 + a3);
                if (hf == 0) a = fma(S.dinv[c], ztp[c], a);
                double* fa = hf ? S.f2 : S.f;
                fa[c] = fma(t, a, fa[c]);
              }
            }
            sp = fma(t, zeta, sp);
          }
          if (tid < q) S.u[tid] = fma(-t, S.r[tid], S.u[tid]);
          up += t;
          MPC_TICK(13);
          if (full) {
            if (q >= QCAP) { status = QPB_MAX_ITER; stop = true; __syncthreads(); break; }  // cannot happen: rows are independent
            if (ci < n) {
              const double val = ztp[ci] * iz;
              for (int k = ck; k < q; k += CKS) {
                double& e = ns_at(S, k, ci);
                e = fma(-S.r[k], val, e);
              }
              if (q % CKS == ck) ns_at(S, q, ci) = val;
            }
            if (tid == 0) {
              S.A[q] = p;
              S.u[q] = up;
              S.slot_of_row[p] = (unsigned char)q;
              S.act[p] = 1;
              S.sv[q] = (unsigned char)vv;
              S.sz[q] = (unsigned char)zz;
              S.scv[q] = cvv;
              S.scz[q] = czz;
            }
            q++;
            __syncthreads();
            MPC_TICK(14);
            break;
          }
          // partial step: slot ks leaves the working set
          for (int k0 = 2 * warp; k0 < q; k0 += 2 * NW) {
            const int j = k0 + hh;
            const bool valid = j < q;
            double a = 0.0;
            if (valid)
              for (int i = hl; i < n; i += 16) a = fma(ns_at(S, j, i), ns_at(S, ks, i), a);
            a = half_sum(a);
            if (valid && hl == 0) S.dd[j] = a;
          }
          __syncthreads();
          const double idl = rcp_fast(S.dd[ks]);
          if (ci < n) {
            const double nu = ns_at(S, ks, ci);
            for (int j = ck; j < q; j += CKS)
              if (j != ks) {
                double& e = ns_at(S, j, ci);
                e = fma(-S.dd[j] * idl, nu, e);
              }
          }
          __syncthreads();
          const int last = q - 1;
          if (ks != last && tid < n) ns_at(S, ks, tid) = ns_at(S, last, tid);
          if (tid == 0) {
            S.act[S.A[ks]] = 0;
            if (ks != last) {
              S.A[ks] = S.A[last];
              S.u[ks] = S.u[last];
              S.slot_of_row[S.A[last]] = (unsigned char)ks;
              S.sv[ks] = S.sv[last];
              S.sz[ks] = S.sz[last];
              S.scv[ks] = S.scv[last];
              S.scz[ks] = S.scz[last];
            }
          }
          q--;
          __syncthreads();
          MPC_TICK(15);
        }
        if (stop) break;
      }
    }

    MPC_TICK(3);
    // ---- output record: U scattered back to the 120 original variables, zeros for swing feet / failures -----
    __syncthreads();
    if (tid < 128) S.rec[tid] = 0.0;
    __syncthreads();
    if (status == QPB_OK && tid < n) {
      const int c = tid / 3, comp = tid - 3 * c;
      S.rec[12 * S.sfk[c] + 3 * S.sff[c] + comp] = S.f[tid] + S.f2[tid];
    }
    if (tid == 0) {
      int* tail = reinterpret_cast<int*>(&S.rec[120]);
      tail[0] = status;
      tail[1] = iters;
    }
    __syncthreads();
    if (tid < 64) {
      double2 v;
      v.x = S.rec[2 * tid];
      v.y = S.rec[2 * tid + 1];
      reinterpret_cast<double2*>(out + rec)[tid] = v;
    }
    rec = (int64_t)gridDim.x + (int64_t)S.ticket;
    __syncthreads();
    MPC_TICK(4);
#ifdef QPB_MPC_PROFILE
    if (tid == 0) atomicAdd(&g_mpc_prof[5], 1ull);
#endif
  }
  // the last CTA out re-arms the work counter (ticket[0] = work counter, ticket[1] = CTAs finished): graph replays and
  // later launches reuse the slot without a memset
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ticket + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {
      ticket[0] = 0ull;
      ticket[1] = 0ull;
      __threadfence();
    }
  }
}

}  // namespace qpbmpc
