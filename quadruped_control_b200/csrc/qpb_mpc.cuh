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

constexpr int NT = 256;  // threads per CTA
constexpr int NW = NT / 32;
constexpr int NH = 10;
constexpr int NV = 120;   // variables before compaction
constexpr int LD = 121;   // odd leading dimension: row- and column-wise sweeps are both bank-conflict free
constexpr int QMAX = 90;  // rows of N* with their own storage; rows 90..119 live in the dead lower triangle of M
constexpr int QCAP = 120; // working-set capacity = the largest possible number of independent rows
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
  double f[NV];
  double nt[NV];
  double zt[NV];
  double w[NV];         // gradient during assembly, N r in the loop
  double E[NH * 12];    // weighted free-response error per step
  double Iinv[NH * 9];
  double Cs[NH], Ss[NH];
  double col[2][128];   // Gauss-Jordan sweep: the pivot column, double-buffered
  double dval[128];     // pivots d_j
  double u[QCAP];
  double r[QCAP];
  double dd[QCAP];
  int A[QCAP];
  unsigned red[NW];
  unsigned char slot_of_row[MROWS];
  unsigned char act[MROWS];
  unsigned char sfk[40], sff[40];
  int ns;
  unsigned ticket;
};

#ifdef QPB_MPC_PROFILE
// developer build only (tools/time_mpc.py): cycles per phase summed over all QPs, [assembly, sweep, start, loop, io, count]
__device__ unsigned long long g_mpc_prof[8];
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

template <int B>
__device__ __forceinline__ void publish_col(const double (&m)[8][8], double* dst, int ty) {
#pragma unroll
  for (int a = 0; a < 8; a++) dst[ty + 16 * a] = m[a][B];
}
template <int A>
__device__ __forceinline__ void assign_row(double (&m)[8][8], const double (&ci)[8], int tx, int j) {
#pragma unroll
  for (int b = 0; b < 8; b++)
    if (tx + 16 * b > j) m[A][b] = -ci[b];
}

// X(i, c) for the inverse factor stored transposed above the diagonal
__device__ __forceinline__ double X_at(const Smem& S, int i, int c) {
  return i > c ? S.M[c * LD + i] : (i == c ? S.dinv[i] : 0.0);
}

// (X v)_i = sum_{c <= i} X[i][c] v_c ; two threads per i (c parity), result valid in the even thread of the pair
// (all 32 lanes must call: the pair is combined with a full-mask shuffle; lanes with !on contribute nothing)
__device__ __forceinline__ double tri_X(const Smem& S, const double* v, int i, int part, bool on) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int c = on ? part : i;
  const double* col = &S.M[i];
  for (; c + 6 < i; c += 8) {
    a0 = fma(col[c * LD], v[c], a0);
    a1 = fma(col[(c + 2) * LD], v[c + 2], a1);
    a2 = fma(col[(c + 4) * LD], v[c + 4], a2);
    a3 = fma(col[(c + 6) * LD], v[c + 6], a3);
  }
  for (; c < i; c += 2) a0 = fma(col[c * LD], v[c], a0);
  double a = (a0 + a1) + (a2 + a3);
  if (on && part == 0) a = fma(S.dinv[i], v[i], a);
  return a + __shfl_xor_sync(FULL, a, 1);
}

// (X^T v)_c = sum_{i >= c} X[i][c] v_i ; two threads per c
__device__ __forceinline__ double tri_XT(const Smem& S, const double* v, int c, int n, int part, bool on) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const double* row = &S.M[c * LD];
  int i = on ? c + 1 + part : n;
  for (; i + 6 < n; i += 8) {
    a0 = fma(row[i], v[i], a0);
    a1 = fma(row[i + 2], v[i + 2], a1);
    a2 = fma(row[i + 4], v[i + 4], a2);
    a3 = fma(row[i + 6], v[i + 6], a3);
  }
  for (; i < n; i += 2) a0 = fma(row[i], v[i], a0);
  double a = (a0 + a1) + (a2 + a3);
  if (on && part == 0) a = fma(S.dinv[c], v[c], a);
  return a + __shfl_xor_sync(FULL, a, 1);
}

// coefficient of pyramid row type t (0..5) on component comp (0..2) of its foot-step
__device__ __forceinline__ double row_coef(int t, int comp, double mu) {
  if (comp == 2) return t < 4 ? mu : (t == 4 ? 1.0 : -1.0);
  if (comp == 0) return t == 0 ? -1.0 : (t == 3 ? 1.0 : 0.0);
  return t == 1 ? -1.0 : (t == 2 ? 1.0 : 0.0);
}

__device__ __forceinline__ double row_slack(int t, double fx, double fy, double fz, const DevParams& P) {
  switch (t) {
    case 0: return fma(P.mu, fz, -fx);
    case 1: return fma(P.mu, fz, -fy);
    case 2: return fma(P.mu, fz, fy);
    case 3: return fma(P.mu, fz, fx);
    case 4: return fz - P.fzmin;
    default: return P.fzmax - fz;
  }
}

__global__ void __launch_bounds__(NT, 1)
mpc_qp_kernel(const __grid_constant__ DevParams P, const qpb_mpc_rec* __restrict__ in, qpb_mpc_out_rec* __restrict__ out,
              int64_t nrec, unsigned long long* __restrict__ ticket, unsigned long long* __restrict__ ticket_clear) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  if (blockIdx.x == 0 && tid == 0) *ticket_clear = 0ull;  // counter of a launch far in the future

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
        for (int s = 0; s < 12; s++) S.E[12 * k + s] = (s < 3 ? P.sLw[s] : P.Lw[s]) * (fr[s] - xr[s]);
      }
      __syncthreads();

      // ---- phase C: attitude response table, sqrt-weighted: ThW[a][k][s], k > step(a) ------------------------
      double* ThW = S.NS;
      for (int idx = tid; idx < n * NH; idx += NT) {
        const int a = idx / NH, k = idx - a * NH, j = S.sfk[a / 3];
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        if (k > j) {
          const double C = S.Cs[k] - S.Cs[j], Sn = S.Ss[k] - S.Ss[j];
          const double g0 = S.G[3 * a], g1 = S.G[3 * a + 1], g2 = S.G[3 * a + 2];
          t0 = P.sLw[0] * dt2 * (C * g0 + Sn * g1);
          t1 = P.sLw[1] * dt2 * (-Sn * g0 + C * g1);
          t2 = P.sLw[2] * dt2 * (double)(k - j) * g2;
        }
        ThW[3 * idx] = t0;
        ThW[3 * idx + 1] = t1;
        ThW[3 * idx + 2] = t2;
      }
      __syncthreads();

      // ---- phase D: H (lower triangle) and the gradient ------------------------------------------------------
      for (int idx = tid; idx < n * n; idx += NT) {
        const int a = idx / n, b = idx - a * n;
        if (b > a) continue;
        const int fa = a / 3, fb = b / 3, ca = a - 3 * fa, cbb = b - 3 * fb;
        const int ja = S.sfk[fa], jb = S.sfk[fb];  // jb <= ja: the compact order is step-major
        const double* Ta = &ThW[3 * NH * a];
        const double* Tb = &ThW[3 * NH * b];
        double t = 0.0, tb1 = 0.0, tb2 = 0.0;
        for (int k = ja + 1; k < NH; k++) {
          t = fma(Ta[3 * k], Tb[3 * k], t);
          tb1 = fma(Ta[3 * k + 1], Tb[3 * k + 1], tb1);
          tb2 = fma(Ta[3 * k + 2], Tb[3 * k + 2], tb2);
        }
        t += tb1 + tb2;
        const double cnt = (double)(NH - ja);
        const double* ga = &S.G[3 * a];
        const double* gb = &S.G[3 * b];
        t += dt2 * cnt * (P.Lw[6] * ga[0] * gb[0] + P.Lw[7] * ga[1] * gb[1] + P.Lw[8] * ga[2] * gb[2]);
        if (ca == cbb) {
          const double s1 = 0.5 * (cnt - 1.0) * cnt, s2 = (cnt - 1.0) * cnt * (2.0 * cnt - 1.0) * (1.0 / 6.0);
          t += (dt * im) * (dt * im) * P.Lw[9 + ca] * cnt;
          t += (dt2 * im) * (dt2 * im) * P.Lw[3 + ca] * (s2 + (double)(ja - jb) * s1);
        }
        if (a == b) t += P.alpha;
        S.M[a * LD + b] = 2.0 * t;
      }
      if (tid < n) {
        const int a = tid, fa = a / 3, ca = a - 3 * fa, ja = S.sfk[fa];
        const double* Ta = &ThW[3 * NH * a];
        const double* ga = &S.G[3 * a];
        double t = 0.0;
        for (int k = ja; k < NH; k++) {
          const double* e = &S.E[12 * k];
          t += Ta[3 * k] * e[0] + Ta[3 * k + 1] * e[1] + Ta[3 * k + 2] * e[2];
          t += (double)(k - ja) * dt2 * im * e[3 + ca];
          t += dt * (ga[0] * e[6] + ga[1] * e[7] + ga[2] * e[8]);
          t += dt * im * e[9 + ca];
        }
        S.w[a] = 2.0 * t;
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
        const int tx = tid & 15, ty = tid >> 4;
        double m[8][8];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int b = 0; b < 8; b++) {
            const int R = ty + 16 * a, C = tx + 16 * b;
            m[a][b] = (R < n && C < n) ? (C <= R ? S.M[R * LD + C] : S.M[C * LD + R]) : 0.0;
          }
        if (tx == 0) publish_col<0>(m, S.col[0], ty);
        __syncthreads();
        for (int j = 0; j < n; j++) {
          const double* cur = S.col[j & 1];
          double* nxt = S.col[(j + 1) & 1];
          const double d = cur[j];
          if (!(d > 0.0)) { status = QPB_BAD_INPUT; break; }  // uniform: every thread reads the same value
          const double rd = rcp_fast(d);
          if (tid == 0) S.dval[j] = d;
          const int bj = j >> 4, tj = j & 15;
          double ci[8], cv[8];
#pragma unroll
          for (int b = 0; b < 8; b++) {
            const int C = tx + 16 * b, R = ty + 16 * b;
            ci[b] = (C > j && C < n) ? cur[C] * rd : 0.0;
            cv[b] = (R != j) ? cur[R] : 0.0;
          }
#pragma unroll
          for (int b = 0; b < 8; b++) {
            if (b < bj || 16 * b >= n) continue;  // uniform: columns already eliminated / beyond the matrix
#pragma unroll
            for (int a = 0; a < 8; a++)
              if (16 * a < n) m[a][b] = fma(-cv[a], ci[b], m[a][b]);
          }
          if (ty == tj) {
            switch (bj) {
              case 0: assign_row<0>(m, ci, tx, j); break;
              case 1: assign_row<1>(m, ci, tx, j); break;
              case 2: assign_row<2>(m, ci, tx, j); break;
              case 3: assign_row<3>(m, ci, tx, j); break;
              case 4: assign_row<4>(m, ci, tx, j); break;
              case 5: assign_row<5>(m, ci, tx, j); break;
              case 6: assign_row<6>(m, ci, tx, j); break;
              default: assign_row<7>(m, ci, tx, j); break;
            }
          }
          const int jn = j + 1;
          if (tx == (jn & 15)) {
            switch (jn >> 4) {
              case 0: publish_col<0>(m, nxt, ty); break;
              case 1: publish_col<1>(m, nxt, ty); break;
              case 2: publish_col<2>(m, nxt, ty); break;
              case 3: publish_col<3>(m, nxt, ty); break;
              case 4: publish_col<4>(m, nxt, ty); break;
              case 5: publish_col<5>(m, nxt, ty); break;
              case 6: publish_col<6>(m, nxt, ty); break;
              default: publish_col<7>(m, nxt, ty); break;
            }
          }
          __syncthreads();
        }
        status = __syncthreads_or(status) ? QPB_BAD_INPUT : QPB_OK;
        if (status == QPB_OK) {
          if (tid < n) S.dinv[tid] = 1.0 / sqrt(S.dval[tid]);
          __syncthreads();
          // X^T above the diagonal: M[c][i] = Xu[i][c] / L_ii
#pragma unroll
          for (int b = 0; b < 8; b++) {
            const int C = tx + 16 * b;
            const double sc = C < n ? S.dinv[C] : 0.0;
#pragma unroll
            for (int a = 0; a < 8; a++) {
              const int R = ty + 16 * a;
              if (C > R && C < n) S.M[R * LD + C] = m[a][b] * sc;
            }
          }
        }
      }
    }

    MPC_TICK(1);
    if (status == QPB_OK && n > 0) {
      __syncthreads();
      // ---- unconstrained minimiser f0 = -X^T X g ----------------------------------------------------------
      const int part = tid & 1;
      const bool on = (tid >> 1) < n;
      const int pi = on ? (tid >> 1) : 0;
      {
        const double y = tri_X(S, S.w, pi, part, on);
        if (on && part == 0) S.zt[pi] = -y;
        __syncthreads();
        const double f0 = tri_XT(S, S.zt, pi, n, part, on);
        if (on && part == 0) S.f[pi] = f0;
        __syncthreads();
      }

      MPC_TICK(2);
      // ---- dual active-set loop ---------------------------------------------------------------------------
      const int m = 2 * n;  // 6 rows per stance foot-step
      const double fzs = 1.0 + fmax(fabs(P.fzmin), fabs(P.fzmax));
      int q = 0;
      for (;;) {
        unsigned key = 0;
        if (tid < m && S.act[tid] == 0) {
          const int c = tid / 6, t = tid - 6 * c;
          const double s = row_slack(t, S.f[3 * c], S.f[3 * c + 1], S.f[3 * c + 2], P);
          const double tol = t < 4 ? 1e-9 : 1e-9 * fzs;
          if (s < -tol) key = ((unsigned)__double2hiint(-s) & 0xffffff00u) | (unsigned)tid;
        }
        key = __reduce_max_sync(FULL, key);
        if (lane == 0) S.red[warp] = key;
        __syncthreads();
        key = S.red[0];
#pragma unroll
        for (int i = 1; i < NW; i++) key = max(key, S.red[i]);
        if (key == 0u) break;  // primal feasible: optimal
        const int p = (int)(key & 0xffu);
        const int pc = p / 6, pt = p - 6 * pc;
        const int zz = 3 * pc + 2, vv = (pt == 0 || pt == 3) ? 3 * pc : ((pt == 1 || pt == 2) ? 3 * pc + 1 : zz);
        const double cvv = pt < 4 ? row_coef(pt, vv - 3 * pc, P.mu) : 0.0, czz = row_coef(pt, 2, P.mu);
        double sp = row_slack(pt, S.f[3 * pc], S.f[3 * pc + 1], S.f[3 * pc + 2], P);
        double up = 0.0;
        bool stop = false;
        for (;;) {  // steps towards row p until it joins the working set
          if (iters >= P.max_iter) { status = QPB_MAX_ITER; stop = true; break; }
          iters++;
          // (1) n~ = X n_p
          if (tid < n) S.nt[tid] = cvv * X_at(S, tid, vv) + czz * X_at(S, tid, zz);
          __syncthreads();
          const double* ztp = S.nt;
          if (q > 0) {
            // (2) r = N* n~
            for (int k = warp; k < q; k += NW) {
              double a = 0.0;
              for (int i = lane; i < n; i += 32) a = fma(ns_at(S, k, i), S.nt[i], a);
              a = warp_sum(a);
              if (lane == 0) S.r[k] = a;
            }
            __syncthreads();
            // (3) w = N r through the sparse rows, then z~ = n~ - X w
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
            const double xw = tri_X(S, S.w, pi, part, on);
            if (on && part == 0) S.zt[pi] = S.nt[pi] - xw;
            __syncthreads();
            ztp = S.zt;
          }
          // (4) zeta = n~.z~, |n~|^2, ratio test -- every warp computes them redundantly (identical results)
          double zeta = 0.0, nn = 0.0;
          for (int i = lane; i < n; i += 32) {
            zeta = fma(S.nt[i], ztp[i], zeta);
            nn = fma(S.nt[i], S.nt[i], nn);
          }
          zeta = warp_sum(zeta);
          nn = warp_sum(nn);
          const bool dep = !(zeta > 1e-13 * nn);
          double t1 = INF;
          int ks = -1;
          for (int k = lane; k < q; k += 32) {
            const double rk = S.r[k];
            if (rk > 0.0) {
              const double ratio = fmax(S.u[k], 0.0) / rk;
              if (ratio < t1) { t1 = ratio; ks = k; }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double ot = __shfl_xor_sync(FULL, t1, o);
            const int ok = __shfl_xor_sync(FULL, ks, o);
            if (ok >= 0 && (ks < 0 || ot < t1 || (ot == t1 && ok < ks))) { t1 = ot; ks = ok; }
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
          const double t2 = dep ? INF : -sp / zeta;
          const bool full = !dep && (!has1 || t2 <= t1);
          const double t = full ? t2 : t1;
          // (5) primal and dual step
          if (!dep) {
            const double df = tri_XT(S, ztp, pi, n, part, on);
            if (on && part == 0) S.f[pi] = fma(t, df, S.f[pi]);
            sp = fma(t, zeta, sp);
          }
          if (tid < q) S.u[tid] = fma(-t, S.r[tid], S.u[tid]);
          up += t;
          if (full) {
            if (q >= QCAP) { status = QPB_MAX_ITER; stop = true; __syncthreads(); break; }  // cannot happen: rows are independent
            const double iz = 1.0 / zeta;
            for (int idx = tid; idx < (q + 1) * n; idx += NT) {
              const int k = idx / n, i = idx - k * n;
              const double val = ztp[i] * iz;
              double& e = ns_at(S, k, i);
              e = k < q ? fma(-S.r[k], val, e) : val;
            }
            if (tid == 0) {
              S.A[q] = p;
              S.u[q] = up;
              S.slot_of_row[p] = (unsigned char)q;
              S.act[p] = 1;
            }
            q++;
            __syncthreads();
            break;
          }
          // partial step: slot ks leaves the working set
          for (int j = warp; j < q; j += NW) {
            double a = 0.0;
            for (int i = lane; i < n; i += 32) a = fma(ns_at(S, j, i), ns_at(S, ks, i), a);
            a = warp_sum(a);
            if (lane == 0) S.dd[j] = a;
          }
          __syncthreads();
          const double idl = 1.0 / S.dd[ks];
          for (int idx = tid; idx < q * n; idx += NT) {
            const int j = idx / n, i = idx - j * n;
            if (j != ks) {
              double& e = ns_at(S, j, i);
              e = fma(-S.dd[j] * idl, ns_at(S, ks, i), e);
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
            }
          }
          q--;
          __syncthreads();
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
      S.rec[12 * S.sfk[c] + 3 * S.sff[c] + comp] = S.f[tid];
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
}

}  // namespace qpbmpc
