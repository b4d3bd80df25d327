// qpb_kernel.cuh -- the balance-controller hot path as one sm_100a kernel family.
//
// One warp solves one robot's QP end to end (north_star): load the 512-B state record with one
// coalesced 16-B load per lane, assemble the 12-variable QP of
// BalanceController::control (reference balance_controller.cpp:98-330), solve it with a dual
// active-set method, then apply the world->body epilogue (balance_controller.cpp:218-232) and
// the Jacobian-transpose torque map (kinematics.cpp:162-188, 218-231), and store 256 B.
//
// Solver (DESIGN.md "Algorithm"): Goldfarb-Idnani dual active set in the WHITENED space
// y = L^T f (Q = L L^T), where the Hessian is the identity.  State per warp:
//   lanes 0..11  : f_i (world-frame force component), row i of the projector P = I - N~ (N~^T N~)^-1 N~^T
//   lanes 16..27 : working-set slot k: constraint id, multiplier u_k, row k of N~* = (N~^T N~)^-1 N~^T
// P and N~* rows live in the same register array M[12], so one DFMA stream updates both.  The 24
// whitened normals n~_j = L^-1 n_j (2-row combinations of J0 = L^-T, friction-pyramid rows have two
// non-zeros) are tabulated in shared memory once per QP.  Working in the whitened space keeps the
// error at ~sqrt(cond(Q)) * eps instead of cond(Q) * eps (cond(Q) ~ 7e5 for W = 1e-5 I).
//
// Inequality rows: variable lane v (leg l = v/3, axis a = v%3) watches rows 2v ("A") and 2v+1 ("B"):
//   a = 0: A: -fx + mu fz >= 0   B:  fx + mu fz >= 0      (rows 0 and 3 of Cf, balance_controller.cpp:278-282)
//   a = 1: A: -fy + mu fz >= 0   B:  fy + mu fz >= 0      (rows 1 and 2)
//   a = 2: A:  fz >= fzmin       B: -fz >= -fzmax         (row 4, both sides)
// The reference's +-1e6 "far" sides (balance_controller.cpp:296-297) are provably inactive when
// 2 mu fzmax <= 1e6, which qpb_create enforces.
//
// No tensor cores: there is no dense contraction here (12x12 FP64 per problem).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qpb200.h"
#include "qpb_stages.h"

namespace qpb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_CTA = 4;
#ifndef QPB_MIN_CTAS_PER_SM
#define QPB_MIN_CTAS_PER_SM 4
#endif
constexpr int LS = 14;  // row stride (doubles) of 12x12 matrices in shared memory: conflict-free 128-bit rows

// Per-warp shared memory.
struct __align__(16) WarpSmem {
  double rec[64];        // staged input record
  double LJ[12 * LS];    // columns of L during factorisation, then rows of J0 = L^-T
  double Nt[24 * LS];    // whitened normals n~_j (12 entries) + |n~_j|^2 in slot 12
  double bz[32];         // broadcast buffer (one slot per lane)
  double bv[16];         // second broadcast buffer
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// read 12 doubles (16-B aligned) from shared memory as six 128-bit loads
__device__ __forceinline__ void lds12(const double* p, double (&v)[12]) {
  const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const double2 t = p2[j];
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void sts12(double* p, const double (&v)[12]) {
  double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
  for (int j = 0; j < 6; j++) p2[j] = make_double2(v[2 * j], v[2 * j + 1]);
}
// 12-term dot product with two independent accumulators.  The kernels are issue-bound: 13 instructions at
// depth 7 measured 2 % faster than 14 at depth 6 (three accumulators).
__device__ __forceinline__ double dot12(const double (&a)[12], const double (&b)[12]) {
  double s0 = a[0] * b[0], s1 = a[1] * b[1];
#pragma unroll
  for (int j = 2; j < 12; j += 2) {
    s0 = fma(a[j], b[j], s0);
    s1 = fma(a[j + 1], b[j + 1], s1);
  }
  return s0 + s1;
}

// Where a record's 64 double slots come from.
struct PackedIO {
  const qpb_state_rec* __restrict__ in;
  qpb_out_rec* __restrict__ out;
};
struct SplitIO {
  const double *Rwb, *Rwb_d, *x, *xdot, *w, *x_d, *xdot_d, *w_d, *feet, *q;
  const uint8_t* contact;
  double *grf, *tau;
  int32_t* status;
};

__device__ __forceinline__ double split_slot(const SplitIO& io, int64_t idx, int m) {
  if (m < 9) return __ldg(io.Rwb + idx * 9 + m);
  if (m < 18) return __ldg(io.Rwb_d + idx * 9 + (m - 9));
  if (m < 21) return __ldg(io.x + idx * 3 + (m - 18));
  if (m < 24) return __ldg(io.xdot + idx * 3 + (m - 21));
  if (m < 27) return __ldg(io.w + idx * 3 + (m - 24));
  if (m < 30) return __ldg(io.x_d + idx * 3 + (m - 27));
  if (m < 33) return __ldg(io.xdot_d + idx * 3 + (m - 30));
  if (m < 36) return __ldg(io.w_d + idx * 3 + (m - 33));
  if (m < 48) return __ldg(io.feet + idx * 12 + (m - 36));
  if (m < 60) return __ldg(io.q + idx * 12 + (m - 48));
  return 0.0;
}

__device__ __forceinline__ double2 load_rec(const PackedIO& io, int64_t idx, int lane) {
  return __ldg(reinterpret_cast<const double2*>(io.in) + idx * 32 + lane);
}
__device__ __forceinline__ double2 load_rec(const SplitIO& io, int64_t idx, int lane) {
  double2 v;
  v.x = split_slot(io, idx, 2 * lane);
  v.y = split_slot(io, idx, 2 * lane + 1);
  if (lane == 30) {
    const uint32_t c = io.contact[idx * 4] | (io.contact[idx * 4 + 1] << 8) | (io.contact[idx * 4 + 2] << 16) |
                       ((uint32_t)io.contact[idx * 4 + 3] << 24);
    v.x = __hiloint2double(0, (int)c);  // little-endian: bytes 480..483 of the packed record
  }
  return v;
}

// wword: the working set at the optimum (bit 2 (3 leg + axis) = row A, the next one = row B -- the same 24-bit word the
// range-space kernels use) with bit 31 set, the next tick's warm start.
__device__ __forceinline__ void store_rec(const PackedIO& io, int64_t idx, int lane, double grf, double tau,
                                          int status, int iters, uint32_t wword) {
  double* o = reinterpret_cast<double*>(io.out + idx);
  if (lane < 12) {
    o[lane] = grf;
    o[12 + lane] = tau;
  } else if (lane < 16) {
    // bytes 192..255: status, iters, working-set word, zero padding -- the whole 256-B record is written
    int4 v = make_int4(0, 0, 0, 0);
    if (lane == 12) { v.x = status; v.y = iters; v.z = (int)wword; }
    reinterpret_cast<int4*>(o + 24)[lane - 12] = v;
  }
}
__device__ __forceinline__ void store_rec(const SplitIO& io, int64_t idx, int lane, double grf, double tau,
                                          int status, int /*iters*/, uint32_t /*wword*/) {
  if (lane < 12) {
    io.grf[idx * 12 + lane] = grf;
    if (io.tau) io.tau[idx * 12 + lane] = tau;
  } else if (lane == 12 && io.status) {
    io.status[idx] = status;
  }
}

// ------------------------------------------------------------------------------------------------
// The kernel.  Persistent: each warp claims records until the batch is exhausted.
// ------------------------------------------------------------------------------------------------
template <class IO>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, QPB_MIN_CTAS_PER_SM)
balance_qp_kernel(const qpb_params* __restrict__ gparams, IO io, int64_t n, unsigned long long* __restrict__ ticket) {
  __shared__ qpb_params P;
  __shared__ WarpSmem wsm[WARPS_PER_CTA];

  {  // stage the controller parameters once per CTA
    const int nw = sizeof(qpb_params) / 8;
    const double* src = reinterpret_cast<const double*>(gparams);
    double* dst = reinterpret_cast<double*>(&P);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  WarpSmem& ws = wsm[wib];
  const int64_t gw = (int64_t)blockIdx.x * WARPS_PER_CTA + wib;
  const int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_CTA;

  // lane roles
  const bool isP = lane < 12;                 // variable / projector row
  const bool isN = lane >= 16 && lane < 28;   // working-set slot
  const int vi = isP ? lane : 0;
  const int leg = vi / 3, ax = vi - 3 * leg;
  const int zl = 3 * leg + 2;  // lane of this leg's fz
  const int axp1 = (ax + 1) % 3, axp2 = (ax + 2) % 3;
  const double mu = P.mu;
  // slack_A = kz fz + kA f - bA,  slack_B = kz fz - kA f - bB  (rows A/B of the table above)
  const double kz = ax < 2 ? mu : 0.0;
  const double kA = ax < 2 ? -1.0 : 1.0;
  const double bA = ax < 2 ? 0.0 : P.fzmin;
  const double bB = ax < 2 ? 0.0 : -P.fzmax;
  const double ntolA = ax < 2 ? -1e-9 : -1e-9 * (1.0 + fabs(P.fzmin));
  const double ntolB = ax < 2 ? -1e-9 : -1e-9 * (1.0 + fabs(P.fzmax));
  const int max_iter = P.max_iter;
  // the inequality row this lane tabulates (lanes 0..23): row j = 2*ov + side on variable ov
  const int cj = lane < 24 ? lane : 0;
  const int c_ov = cj >> 1, c_leg = c_ov / 3, c_ax = c_ov - 3 * c_leg;
  const int c_ib = 3 * c_leg + 2;
  const double c_ca = (((cj & 1) == 0) == (c_ax < 2)) ? -1.0 : 1.0;
  const double c_cb = (c_ax < 2) ? mu : 0.0;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  // Work distribution: warp w starts on record w; further records are claimed from a global ticket
  // counter (iteration counts vary a lot between QPs, so a static stride leaves a long tail).  The
  // ticket for the next record is requested at the start of the current one, hiding the atomic.
  int64_t idx = gw;
  while (idx < n) {
    unsigned long long next_ticket = 0;
    if (lane == 0) next_ticket = atomicAdd(ticket, 1ULL);
    // ---- load + stage ---------------------------------------------------------------------------
    const double2 v = load_rec(io, idx, lane);
    bool ok = (lane >= 30) || (isfinite(v.x) && isfinite(v.y));
    __syncwarp();
    reinterpret_cast<double2*>(ws.rec)[lane] = v;
    __syncwarp();
    ok = __all_sync(FULL, ok);
    const uint32_t cbytes = reinterpret_cast<const uint32_t*>(ws.rec + 60)[0];
    const uint32_t smask = ((cbytes & 0xffu) ? 1u : 0u) | ((cbytes & 0xff00u) ? 2u : 0u) |
                           ((cbytes & 0xff0000u) ? 4u : 0u) | ((cbytes & 0xff000000u) ? 8u : 0u);
    const bool stance = isP && ((smask >> leg) & 1u);

    int status = QPB_OK, iters = 0;
    uint32_t wset = 0;  // working set the solve ended on
    double x = 0.0;  // f_w component (lanes 0..11)

    if (ok) {
      const double* R = ws.rec;
      // ---- PD target, balance_controller.cpp:126-139 (uniform across lanes) ---------------------
      double b6[6];
      pd_rhs(P, ws.rec, b6);

      // ---- lever arms r_leg = R p_leg (:245-248); lane i holds component ax of leg ----------------
      const double ri = R[3 * ax] * ws.rec[36 + 3 * leg] + R[3 * ax + 1] * ws.rec[37 + 3 * leg] +
                        R[3 * ax + 2] * ws.rec[38 + 3 * leg];
      ws.bz[lane] = ri;
      __syncwarp();
      double rr[12];
      lds12(ws.bz, rr);
      // column i of A = [e_ax ; column ax of skew(r_leg)], rigid3d.cpp:61-74
      const double g_up = ws.bz[3 * leg + axp2];
      const double g_dn = -ws.bz[3 * leg + axp1];
      double ang[3];
#pragma unroll
      for (int k = 0; k < 3; k++) ang[k] = (k == axp1) ? g_up : ((k == axp2) ? g_dn : 0.0);
      // v = S a_i
      double sv[6];
#pragma unroll
      for (int m = 0; m < 6; m++)
        sv[m] = P.S[6 * m + ax] + P.S[6 * m + 3] * ang[0] + P.S[6 * m + 4] * ang[1] + P.S[6 * m + 5] * ang[2];
      // c_i = -2 a_i^T S b, :153
      double ci = 0.0;
#pragma unroll
      for (int m = 0; m < 6; m++) ci = fma(sv[m], b6[m], ci);
      ci = stance ? -2.0 * ci : 0.0;
      // row i of Q = 2 (A^T S A + W), :152; swing variables are decoupled (identity row)
      double Qr[12];
#pragma unroll
      for (int j = 0; j < 12; j++) {
        const int lj = j / 3, aj = j % 3, aj1 = (aj + 1) % 3, aj2 = (aj + 2) % 3;
        double val = sv[aj] + sv[3 + aj1] * rr[3 * lj + aj2] - sv[3 + aj2] * rr[3 * lj + aj1];
        val = 2.0 * (val + P.W[12 * vi + j]);
        const bool both = stance && ((smask >> lj) & 1u);
        Qr[j] = both ? val : ((j == vi) ? 1.0 : 0.0);
      }
      __syncwarp();  // bz is reused below

      // ---- Cholesky Q = L L^T, right-looking; column k of L is published to shared memory.  The
      //      gradient rides along as an extra column, which yields y0 = -L^-1 c by forward substitution.
      double rsd[12];
      double cy = -ci;
#pragma unroll
      for (int k = 0; k < 12; k++) {
        const double d = shfl_d(Qr[k], k);
        ok = ok && (d > 0.0) && (d < 1e300);
        const double rs = rsqrt_fast(d);
        rsd[k] = rs;
        const double lik = (lane >= k) ? Qr[k] * rs : 0.0;  // L[lane][k]
        if (isP) ws.LJ[k * LS + lane] = lik;
        if (lane == k) {
          const double yk = cy * rs;
          ws.LJ[k * LS + 12] = yk;
          ws.bv[k] = yk;
        }
        __syncwarp();
#pragma unroll
        for (int j = k + 1; j < 12; j++) Qr[j] = fma(-lik, ws.LJ[k * LS + j], Qr[j]);
        cy = fma(-lik, ws.LJ[k * LS + 12], cy);
      }
      // ---- T = L^-1 by columns: lane j holds column j of T = row j of J0 = L^-T -----------------
      double J0r[12];
#pragma unroll
      for (int m = 0; m < 12; m++) J0r[m] = (lane == m) ? 1.0 : 0.0;
#pragma unroll
      for (int l = 0; l < 12; l++) {
        J0r[l] *= rsd[l];
#pragma unroll
        for (int m = l + 1; m < 12; m++) J0r[m] = fma(-ws.LJ[l * LS + m], J0r[l], J0r[m]);
      }
      {
        double y0[12];
        lds12(ws.bv, y0);
        x = dot12(J0r, y0);  // unconstrained minimiser f0 = J0 y0 (lanes 0..11)
      }
      __syncwarp();  // everyone is done reading L
      if (isP) sts12(ws.LJ + LS * lane, J0r);  // J0 rows
      __syncwarp();
      // ---- whitened normals n~_j = J0^T n_j and their squared norms (lanes 0..23) ----------------
      {
        double ra[12], rb[12];
        lds12(ws.LJ + LS * c_ov, ra);
        lds12(ws.LJ + LS * c_ib, rb);
#pragma unroll
        for (int m = 0; m < 12; m++) ra[m] = fma(c_cb, rb[m], c_ca * ra[m]);
        if (lane < 24) {
          sts12(ws.Nt + LS * lane, ra);
          ws.Nt[LS * lane + 12] = dot12(ra, ra);
        }
      }
      __syncwarp();

      // ---- dual active set ---------------------------------------------------------------------
      double M[12];
#pragma unroll
      for (int j = 0; j < 12; j++) M[j] = (isP && j == lane) ? 1.0 : 0.0;
      double u = 0.0;
      int cons = -1;
      uint32_t active = 0, ignore = 0;
      int p = -1;
      double up = 0.0;

      if (ok) {
        for (;;) {
          __syncwarp();
          // (1) slacks of the two rows this lane watches:  kz fz +- kA f - b
          const double xz = shfl_d(x, zl);
          const double base = kz * xz;
          const double sA = fma(kA, x, base - bA);
          const double sB = fma(-kA, x, base - bB);
          const uint32_t act2 = (active | ignore) >> ((2 * lane) & 31);
          const bool vA = stance && !(act2 & 1u) && (sA < ntolA);
          const bool vB = stance && !(act2 & 2u) && (sB < ntolB);
          // key: most negative slack wins (sign bit set => larger magnitude = larger unsigned); low 5 bits = row
          const uint32_t keyA = vA ? (((uint32_t)__double2hiint(sA) & ~31u) | (uint32_t)(2 * lane)) : 0u;
          const uint32_t keyB = vB ? (((uint32_t)__double2hiint(sB) & ~31u) | (uint32_t)(2 * lane + 1)) : 0u;
          const uint32_t kmax = __reduce_max_sync(FULL, max(keyA, keyB));
          const bool fresh = p < 0;
          if ((fresh && kmax == 0u) || iters >= max_iter) {  // primal feasible (optimal) or out of iterations
            if (!(fresh && kmax == 0u)) status = QPB_MAX_ITER;
            break;
          }
          if (fresh) {
            p = (int)(kmax & 31u);
            up = 0.0;
          }
          iters++;
          const double sp = shfl_d((p & 1) ? sB : sA, p >> 1);

          // (2) z~ = P n~ (lanes 0..11), r = N~* n~ (lanes 16..27)
          const double* ntp = ws.Nt + LS * p;
          double mv;
          {
            double nt[12];
            lds12(ntp, nt);
            mv = dot12(M, nt);
          }
          const double nn = ntp[12];
          ws.bz[lane] = mv;
          __syncwarp();
          double zt[12];
          lds12(ws.bz, zt);
          // (3) lanes 0..11: dx = J0 z~ ; other lanes: zeta = n~^T z~ (exactly annihilates n~ in the update)
          double acc;
          {
            double a[12];
            lds12(isP ? (ws.LJ + LS * lane) : ntp, a);
            acc = dot12(a, zt);
          }
          const double zeta = shfl_d(acc, 16);
          const bool dep = !(zeta > 1e-13 * nn);  // n~ in the span of the working set
          const double izeta = rcp_fast(zeta);
          const double t2 = -sp * izeta;  // > 0: row p is violated and zeta > 0
          // (4) dual step bound: min u_k / r_k over r_k > 0 (exact argmin via two integer reductions)
          const bool cand = cons >= 0 && mv > 0.0;  // cons >= 0 only on slot lanes
          const double uu = (__double2hiint(u) < 0) ? 0.0 : u;  // rounding can leave -1e-17
          const double ratio = uu * rcp_fast(mv);
          const uint32_t rhi = cand ? (uint32_t)__double2hiint(ratio) : 0x7ff00000u;
          const uint32_t mhi = __reduce_min_sync(FULL, rhi);
          const bool c2 = cand && rhi == mhi;
          const uint32_t rlo = c2 ? (uint32_t)__double2loint(ratio) : 0xffffffffu;
          const uint32_t mlo = __reduce_min_sync(FULL, rlo);
          const uint32_t wb = __ballot_sync(FULL, c2 && rlo == mlo);
          const bool has1 = wb != 0u;
          const int kl = __ffs(wb) - 1;  // -1: no blocking row
          const double t1 = shfl_d(ratio, kl & 31);
          if (dep && !has1) {
            // Row p lies in the span of the working set and no multiplier can give way.  The feasible set is
            // never empty (qpb_create), so this is rounding making the twin of an active row look violated
            // (e.g. fzmin == fzmax): the row holds to rounding, set it aside.
            if (sp < -1e-6 * (1.0 + fmax(fabs(P.fzmin), fabs(P.fzmax)))) {  // not a rounding artefact: give up loudly
              status = QPB_BAD_INPUT;
              break;
            }
            ignore |= 1u << p;
            p = -1;
            continue;
          }
          const bool full = !dep && (!has1 || t2 <= t1);
          const double t = full ? t2 : t1;
          // (5) step
          x = fma(dep ? 0.0 : t, acc, x);  // meaningful on lanes 0..11
          u = fma(-t, mv, u);  // free slots and idle lanes have M = 0, hence mv = 0
          up += t;
          double coef;
          if (full) {
            // (6a) full step: row p enters the working set;  M -= (M n~) z~^T / zeta
            const uint32_t fb = __ballot_sync(FULL, isN && cons < 0);
            const int ql = __ffs(fb) - 1;
            coef = mv * izeta;  // zero on free slots / idle lanes (their M rows are zero)
            if (lane == ql) { coef = -izeta; cons = p; u = up; }
            active |= 1u << p;
            p = -1;
          } else {
            // (6b) partial step: the blocking row (slot kl) leaves;  M += / -= (..) nu^T / |nu|^2
            if (lane == kl) sts12(ws.bv, M);
            __syncwarp();
            lds12(ws.bv, zt);  // zt now holds nu = row kl of N~*
            const double gam = dot12(M, zt);
            const double idelta = rcp_fast(shfl_d(gam, kl));  // 1 / |nu|^2
            const int cdrop = __shfl_sync(FULL, cons, kl);
            coef = 0.0;
            if (isP) coef = -ws.bv[lane] * idelta;
            else if (cons >= 0) coef = gam * idelta;
            if (lane == kl) { coef = 1.0; cons = -1; u = 0.0; }
            active &= ~(1u << cdrop);
          }
#pragma unroll
          for (int j = 0; j < 12; j++) M[j] = fma(-coef, zt[j], M[j]);
        }
        wset = active;
      } else {
        status = QPB_BAD_INPUT;
      }
    } else {
      status = QPB_BAD_INPUT;
    }

    // ---- epilogue: body-frame GRF (:218-232) and tau = J^T f (kinematics.cpp:162-188, 218-231) ---
    const bool good = (status == QPB_OK);
    const double fw = (good && stance) ? x : 0.0;
    const double f0 = shfl_d(fw, 3 * leg), f1 = shfl_d(fw, 3 * leg + 1), f2 = shfl_d(fw, 3 * leg + 2);
    double fb = -1.0 * (ws.rec[ax] * f0 + ws.rec[3 + ax] * f1 + ws.rec[6 + ax] * f2);  // -(R^T f)_ax
    if (!(good && stance)) fb = 0.0;
    // angle handled by this lane: t1, t2, t2+t3
    const double qa = ws.rec[48 + 3 * leg + ax] + ((ax == 2) ? ws.rec[48 + 3 * leg + 1] : 0.0);
    double sn, cs;
    sincos(qa, &sn, &cs);
    const double s1 = shfl_d(sn, 3 * leg), c1 = shfl_d(cs, 3 * leg);
    const double s2 = shfl_d(sn, 3 * leg + 1), c2 = shfl_d(cs, 3 * leg + 1);
    const double s23 = shfl_d(sn, 3 * leg + 2), c23 = shfl_d(cs, 3 * leg + 2);
    const double fbx = shfl_d(fb, 3 * leg), fby = shfl_d(fb, 3 * leg + 1), fbz = shfl_d(fb, 3 * leg + 2);
    const double l1 = P.link[3 * leg], l2 = P.link[3 * leg + 1], l3 = P.link[3 * leg + 2];
    double Jx, Jy, Jz;  // column ax of the leg Jacobian
    leg_jacobian_col(ax, l1, l2, l3, s1, c1, s2, c2, s23, c23, Jx, Jy, Jz);
    double tau = Jx * fbx + Jy * fby + Jz * fbz;
    if (P.clamp_tau) tau = fmin(fmax(tau, P.tau_min), P.tau_max);  // commander_node.cpp:526
    if (!(good && stance)) tau = 0.0;
    store_rec(io, idx, lane, fb, tau, status, iters, wset | 0x80000000u);
    idx = nwarps + (int64_t)__shfl_sync(FULL, next_ticket, 0);
  }
  // The last CTA out re-arms the work counter: the launch is self-contained, so the same counter slot serves graph
  // replays and later launches without a memset (ticket[0] = work counter, ticket[1] = CTAs finished).
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL) {
      ticket[0] = 0ULL;
      ticket[1] = 0ULL;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stand-alone kinematics kernels (one thread per leg).
// ------------------------------------------------------------------------------------------------
__global__ void jt_kernel(const qpb_params* __restrict__ P, const double* __restrict__ q,
                          const double* __restrict__ f, const uint8_t* __restrict__ contact,
                          double* __restrict__ tau, int64_t nlegs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlegs) return;
  const int leg = (int)(i & 3);
  const double l1 = P->link[3 * leg], l2 = P->link[3 * leg + 1], l3 = P->link[3 * leg + 2];
  const double t1 = q[3 * i], t2 = q[3 * i + 1], t3 = q[3 * i + 2];
  double s1, c1, s2, c2, s23, c23;
  sincos(t1, &s1, &c1);
  sincos(t2, &s2, &c2);
  sincos(t2 + t3, &s23, &c23);
  const bool st = contact ? (contact[i] != 0) : true;
  const double fx = f[3 * i], fy = f[3 * i + 1], fz = f[3 * i + 2];
  double o[3];
  leg_jt(l1, l2, l3, s1, c1, s2, c2, s23, c23, fx, fy, fz, o);
  double o0 = o[0], o1 = o[1], o2 = o[2];
  if (P->clamp_tau) {
    o0 = fmin(fmax(o0, P->tau_min), P->tau_max);
    o1 = fmin(fmax(o1, P->tau_min), P->tau_max);
    o2 = fmin(fmax(o2, P->tau_min), P->tau_max);
  }
  tau[3 * i] = st ? o0 : 0.0;
  tau[3 * i + 1] = st ? o1 : 0.0;
  tau[3 * i + 2] = st ? o2 : 0.0;
}

__global__ void fk_kernel(const qpb_params* __restrict__ P, const double* __restrict__ q,
                          double* __restrict__ feet, int64_t nlegs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlegs) return;
  const int leg = (int)(i & 3);
  const double l1 = P->link[3 * leg], l2 = P->link[3 * leg + 1], l3 = P->link[3 * leg + 2];
  const double t1 = q[3 * i], t2 = q[3 * i + 1], t3 = q[3 * i + 2];
  double s1, c1, s2, c2, s23, c23;
  sincos(t1, &s1, &c1);
  sincos(t2, &s2, &c2);
  sincos(t2 + t3, &s23, &c23);
  feet[3 * i] = l2 * s2 + l3 * s23 + P->hip_offset[3 * leg];
  feet[3 * i + 1] = l1 * c1 - l2 * s1 * c2 - l3 * s1 * c23 + P->hip_offset[3 * leg + 1];
  feet[3 * i + 2] = l1 * s1 + l2 * c1 * c2 + l3 * c1 * c23 + P->hip_offset[3 * leg + 2];
}

}  // namespace qpb
