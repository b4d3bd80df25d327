// qpb_kernel16.cuh -- half-warp variant of the balance kernel: TWO QPs per warp, 16 lanes each.
//
// Same algorithm and arithmetic as balance_qp_kernel (qpb_kernel.cuh).  What
// changes is the mapping: the 12 variable lanes of a half-warp own BOTH the projector row (Mp) and the
// working-set slot / pseudo-inverse row (Mn) of their QP, lane 12 of the half computes zeta, and the
// two halves run in lock step, so every bookkeeping instruction (slack tests, decode, step logic,
// shared-memory broadcasts) is paid once for two problems.  Per-half reductions are pairs of warp-wide
// REDUX with the other half's lanes neutralised.  A half that finishes early idles until its partner is done.
#pragma once

#include "qpb_kernel.cuh"

namespace qpb {

// scalars of the solver loop, passed by value as a kernel argument (constant bank)
struct LoopConsts {
  double mu, fzmin, fzmax;
  double ntol_z;  // violation tolerance of the fz rows: -1e-9 (1 + max(|fzmin|, |fzmax|))
  int max_iter;
};

struct __align__(16) HalfSmem {
  double rec[64];        // staged input record
  double LJ[12 * LS];    // columns of L during factorisation, then rows of J0 = L^-T
  double Nt[24 * LS];    // whitened normals n~_j (12 entries) + |n~_j|^2 in slot 12
  double bz[16];         // broadcast buffer (one slot per lane of the half)
  double bv[16];         // second broadcast buffer
};

__device__ __forceinline__ double shfl16(double v, int src) { return __shfl_sync(FULL, v, src, 16); }

// Per-half reductions.  Measured on config 2: 4-step shuffle butterflies 1.82e8 QP/s; REDUX with a half-warp
// member mask 1.69e8 (the compiler serialises the two masks); two full-mask REDUX with the other half
// neutralised 1.92e8 -- the default.
#ifdef QPB_HALF_BUTTERFLY
__device__ __forceinline__ uint32_t half_max_u32(uint32_t v, bool) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o, 16));
  return v;
}
__device__ __forceinline__ double half_min_f64(double v, bool) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(FULL, v, o, 16);
    v = (w < v) ? w : v;
  }
  return v;
}
#else
// Two full-mask REDUX per reduction, one per half with the other half's lanes neutralised: the results land in
// uniform registers and the two are independent, so the chain is ~1 REDUX deep instead of 4 dependent shuffles.
__device__ __forceinline__ uint32_t half_max_u32(uint32_t v, bool upper) {
  const uint32_t a = __reduce_max_sync(FULL, upper ? 0u : v);
  const uint32_t b = __reduce_max_sync(FULL, upper ? v : 0u);
  return upper ? b : a;
}
// exact minimum of non-negative doubles (or +inf): their bit patterns order like unsigned integers
__device__ __forceinline__ double half_min_f64(double v, bool upper) {
  const uint32_t hi = (uint32_t)__double2hiint(v);
  const uint32_t ah = __reduce_min_sync(FULL, upper ? 0xffffffffu : hi);
  const uint32_t bh = __reduce_min_sync(FULL, upper ? hi : 0xffffffffu);
  const uint32_t mh = upper ? bh : ah;
  const uint32_t lo = (hi == mh) ? (uint32_t)__double2loint(v) : 0xffffffffu;
  const uint32_t al = __reduce_min_sync(FULL, upper ? 0xffffffffu : lo);
  const uint32_t bl = __reduce_min_sync(FULL, upper ? lo : 0xffffffffu);
  return __hiloint2double((int)mh, (int)(upper ? bl : al));
}
#endif

// 512-B record of this half: lane l holds slots {2l, 2l+1} (a) and {32+2l, 33+2l} (b)
__device__ __forceinline__ void load_rec16(const PackedIO& io, int64_t rec, int l, double2& a, double2& b) {
  const double2* p = reinterpret_cast<const double2*>(io.in) + rec * 32;
  a = __ldg(p + l);
  b = __ldg(p + 16 + l);
}
__device__ __forceinline__ void load_rec16(const SplitIO& io, int64_t rec, int l, double2& a, double2& b) {
  a.x = split_slot(io, rec, 2 * l);
  a.y = split_slot(io, rec, 2 * l + 1);
  b.x = split_slot(io, rec, 32 + 2 * l);
  b.y = split_slot(io, rec, 33 + 2 * l);
  if (l == 14) {
    const uint32_t c = io.contact[rec * 4] | (io.contact[rec * 4 + 1] << 8) | (io.contact[rec * 4 + 2] << 16) |
                       ((uint32_t)io.contact[rec * 4 + 3] << 24);
    b.x = __hiloint2double(0, (int)c);
  }
}

template <class IO>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, QPB_MIN_CTAS_PER_SM)
balance_qp_kernel16(const qpb_params* __restrict__ gparams, IO io, int64_t n, unsigned long long* __restrict__ ticket,
                    const LoopConsts kc) {
  __shared__ qpb_params P;
  __shared__ HalfSmem hsm[WARPS_PER_CTA * 2];

  {
    const int nw = sizeof(qpb_params) / 8;
    const double* src = reinterpret_cast<const double*>(gparams);
    double* dst = reinterpret_cast<double*>(&P);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  // warp index through a warp-wide reduction: the result lives in a uniform register, so the base address of this
  // warp's shared-memory block is formed on the uniform datapath instead of being rebuilt from tid in every iteration
  const int wib = (int)__reduce_max_sync(FULL, threadIdx.x >> 5);
  const int l = lane & 15;         // lane within the half
  const int hb = lane & 16;        // first lane of this half
  HalfSmem& hs = hsm[2 * wib + (lane >> 4)];
  // 32-bit work indices (qpb_api.cu bounds n per launch): fewer live registers in the solver loop
  const uint32_t gw = blockIdx.x * WARPS_PER_CTA + wib;
  const uint32_t nwarps = gridDim.x * WARPS_PER_CTA;
  const uint32_t npairs = (uint32_t)((n + 1) >> 1);

  const bool isP = l < 12;  // variable lane: projector row, working-set slot, pseudo-inverse row
  const int vi = isP ? l : 0;
  const int leg = vi / 3, ax = vi - 3 * leg;
  const int zl = 3 * leg + 2;
  const int axp1 = (ax + 1) % 3, axp2 = (ax + 2) % 3;
  // kc: the scalars the solver loop uses, passed as kernel arguments: they sit in the constant bank and are used as
  // instruction operands, costing neither registers nor shared-memory loads
  const double mu = kc.mu;
  const double kz = ax < 2 ? mu : 0.0;
  const int sgnA = ax < 2 ? (int)0x80000000 : 0;  // sign applied to f in row A (row B uses the opposite)
  const bool lat = ax < 2;  // this lane watches the two pyramid rows of a lateral variable (else the two fz rows)
  // a row is violated when its slack is below -1e-9 (1 + |bound|); one value per lane (the larger bound of its two rows)

  uint32_t pair = gw;
  while (pair < npairs) {
    uint32_t next_ticket = 0;
    if (lane == 0) next_ticket = (uint32_t)atomicAdd(ticket, 1ULL);
    const int64_t rec = 2 * (int64_t)pair + (lane >> 4);
    const bool have = rec < n;  // the last pair may be half empty
    // ---- load + stage ---------------------------------------------------------------------------
    double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
    if (have) load_rec16(io, rec, l, va, vb);
    bool okl = isfinite(va.x) && isfinite(va.y) && ((l >= 14) || (isfinite(vb.x) && isfinite(vb.y)));
    __syncwarp();
    reinterpret_cast<double2*>(hs.rec)[l] = va;
    reinterpret_cast<double2*>(hs.rec)[16 + l] = vb;
    __syncwarp();
    const uint32_t okb = __ballot_sync(FULL, okl);
    bool ok = have && (((okb >> hb) & 0xffffu) == 0xffffu);
    const uint32_t cbytes = reinterpret_cast<const uint32_t*>(hs.rec + 60)[0];
    const uint32_t smask = ((cbytes & 0xffu) ? 1u : 0u) | ((cbytes & 0xff00u) ? 2u : 0u) |
                           ((cbytes & 0xff0000u) ? 4u : 0u) | ((cbytes & 0xff000000u) ? 8u : 0u);
    const bool stance = isP && ((smask >> leg) & 1u);

    int status = QPB_OK, iters = 0;
    double x = 0.0;

    const double* R = hs.rec;
    // ---- PD target + dynamics right-hand side (qpb_stages.h; uniform across the half) ----------------
    double b6[6];
    pd_rhs(P, hs.rec, b6);

    // ---- lever arms r_leg = R p_leg (:245-248) ----------------------------------------------------
    const double ri = R[3 * ax] * hs.rec[36 + 3 * leg] + R[3 * ax + 1] * hs.rec[37 + 3 * leg] +
                      R[3 * ax + 2] * hs.rec[38 + 3 * leg];
    hs.bz[l] = ri;
    __syncwarp();
    double rr[12];
    lds12(hs.bz, rr);
    const double g_up = hs.bz[3 * leg + axp2];
    const double g_dn = -hs.bz[3 * leg + axp1];
    double ang[3];
#pragma unroll
    for (int k = 0; k < 3; k++) ang[k] = (k == axp1) ? g_up : ((k == axp2) ? g_dn : 0.0);
    double sv[6];
#pragma unroll
    for (int m = 0; m < 6; m++)
      sv[m] = P.S[6 * m + ax] + P.S[6 * m + 3] * ang[0] + P.S[6 * m + 4] * ang[1] + P.S[6 * m + 5] * ang[2];
    double ci = 0.0;
#pragma unroll
    for (int m = 0; m < 6; m++) ci = fma(sv[m], b6[m], ci);
    ci = stance ? -2.0 * ci : 0.0;  // :153
    double Qr[12];                  // row vi of Q = 2 (A^T S A + W), :152; swing variables decoupled
#pragma unroll
    for (int j = 0; j < 12; j++) {
      const int lj = j / 3, aj = j % 3, aj1 = (aj + 1) % 3, aj2 = (aj + 2) % 3;
      double val = sv[aj] + sv[3 + aj1] * rr[3 * lj + aj2] - sv[3 + aj2] * rr[3 * lj + aj1];
      val = 2.0 * (val + P.W[12 * vi + j]);
      const bool both = stance && ((smask >> lj) & 1u);
      Qr[j] = both ? val : ((j == vi) ? 1.0 : 0.0);
    }
    __syncwarp();

    // ---- Cholesky (right-looking) with the gradient as an extra column -> y0 = -L^-1 c ------------
    double rsd[12];
    double cy = -ci;
#pragma unroll
    for (int k = 0; k < 12; k++) {
      const double d = shfl16(Qr[k], k);
      ok = ok && (d > 0.0) && (d < 1e300);
      const double rs = rsqrt_fast(d);
      rsd[k] = rs;
      const double lik = (l >= k) ? Qr[k] * rs : 0.0;  // L[l][k]
      if (isP) hs.LJ[k * LS + l] = lik;
      if (l == k) {
        const double yk = cy * rs;
        hs.LJ[k * LS + 12] = yk;
        hs.bv[k] = yk;
      }
      __syncwarp();
#pragma unroll
      for (int j = k + 1; j < 12; j++) Qr[j] = fma(-lik, hs.LJ[k * LS + j], Qr[j]);
      cy = fma(-lik, hs.LJ[k * LS + 12], cy);
    }
    // ---- T = L^-1 by columns: lane j holds row j of J0 = L^-T --------------------------------------
    double J0r[12];
#pragma unroll
    for (int m = 0; m < 12; m++) J0r[m] = (l == m) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < 12; c++) {
      J0r[c] *= rsd[c];
#pragma unroll
      for (int m = c + 1; m < 12; m++) J0r[m] = fma(-hs.LJ[c * LS + m], J0r[c], J0r[m]);
    }
    {
      double y0[12];
      lds12(hs.bv, y0);
      x = dot12(J0r, y0);  // unconstrained minimiser f0 = J0 y0
    }
    __syncwarp();
    if (isP) sts12(hs.LJ + LS * l, J0r);
    __syncwarp();
    // ---- whitened normals of the two rows this variable lane watches --------------------------------
    {
      double ra[12], rb[12];
      lds12(hs.LJ + LS * vi, ra);
      lds12(hs.LJ + LS * zl, rb);
      double na[12], nb[12];
      // rows A/B:  -/+ J0[v] + mu J0[z]  (ax < 2);   +/- J0[z]  (ax == 2)
      const double kA = ax < 2 ? -1.0 : 1.0;
#pragma unroll
      for (int m = 0; m < 12; m++) {
        const double base = kz * rb[m];
        na[m] = fma(kA, ra[m], base);
        nb[m] = fma(-kA, ra[m], base);
      }
      if (isP) {
        sts12(hs.Nt + LS * (2 * l), na);
        sts12(hs.Nt + LS * (2 * l + 1), nb);
        hs.Nt[LS * (2 * l) + 12] = dot12(na, na);
        hs.Nt[LS * (2 * l + 1) + 12] = dot12(nb, nb);
      }
    }
    __syncwarp();

    // ---- dual active set, both halves in lock step --------------------------------------------------
    double Mp[12], Mn[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
      Mp[j] = (isP && j == l) ? 1.0 : 0.0;
      Mn[j] = 0.0;
    }
    double u = 0.0;
    int cons = -1;
    uint32_t active = 0, ignore = 0;
    int p = -1;
    double up = 0.0;
    bool done = !ok;
    if (!ok) status = QPB_BAD_INPUT;

    for (;;) {
      __syncwarp();
      // (1) slacks of the two rows this lane watches
      const double xz = shfl16(x, zl);
      const double base = kz * xz;
      const double xs = __hiloint2double(__double2hiint(x) ^ sgnA, __double2loint(x));  // -x (pyramid rows) or +x (fz rows)
      const double sA = (base - (lat ? 0.0 : kc.fzmin)) + xs;
      const double sB = (base + (lat ? 0.0 : kc.fzmax)) - xs;
      const double ntol = lat ? -1e-9 : kc.ntol_z;
      const uint32_t act2 = (active | ignore) >> ((2 * l) & 31);
      const bool vA = stance && !(act2 & 1u) && (sA < ntol);
      const bool vB = stance && !(act2 & 2u) && (sB < ntol);
      const uint32_t keyA = vA ? (((uint32_t)__double2hiint(sA) & ~31u) | (uint32_t)(2 * l)) : 0u;
      const uint32_t keyB = vB ? (((uint32_t)__double2hiint(sB) & ~31u) | (uint32_t)(2 * l + 1)) : 0u;
      const uint32_t kmax = half_max_u32(max(keyA, keyB), hb != 0);
      const bool fresh = p < 0;
      if (!done && ((fresh && kmax == 0u) || iters >= kc.max_iter)) {
        if (!(fresh && kmax == 0u)) status = QPB_MAX_ITER;
        done = true;
      }
      if (__all_sync(FULL, done)) break;
      if (!done && fresh) {
        p = (int)(kmax & 31u);
        up = 0.0;
      }
      if (!done) iters++;
      const int pp = p < 0 ? 0 : p;  // finished halves keep executing on row 0 with a zero step
      const double sp = shfl16((pp & 1) ? sB : sA, pp >> 1);

      // (2) z~ = P n~, r = N~* n~
      const double* ntp = hs.Nt + LS * pp;
      double mvP, mvN;
      {
        double nt[12];
        lds12(ntp, nt);
        mvP = dot12(Mp, nt);
        mvN = dot12(Mn, nt);
      }
      const double nn = ntp[12];
      hs.bz[l] = mvP;
      __syncwarp();
      double zt[12];
      lds12(hs.bz, zt);
      // (3) variable lanes: dx = J0 z~ ; lane 12 of the half: zeta = n~^T z~
      double acc;
      {
        double a[12];
        lds12(isP ? (hs.LJ + LS * l) : ntp, a);
        acc = dot12(a, zt);
      }
      const double zeta = shfl16(acc, 12);
      const bool dep = !(zeta > 1e-13 * nn);
      const double izeta = rcp_fast(zeta);
      const double t2 = -sp * izeta;
      // (4) dual step bound
      const bool cand = cons >= 0 && mvN > 0.0;
      const double uu = (__double2hiint(u) < 0) ? 0.0 : u;
      const double INF = __longlong_as_double(0x7ff0000000000000LL);
      const double ratio = cand ? uu * rcp_fast(mvN) : INF;
      const double t1 = half_min_f64(ratio, hb != 0);
      const bool has1 = t1 < INF;
      const uint32_t wb = (__ballot_sync(FULL, cand && ratio == t1) >> hb) & 0xffffu;
      const int kl = __ffs(wb) - 1;  // lane (within the half) of the blocking slot, -1 if none
      // Row p lies in the span of the working set and no multiplier can give way: with a non-empty feasible
      // set (qpb_create guarantees one) this only happens when rounding makes the twin of an active row look
      // violated (e.g. fzmin == fzmax).  The row is satisfied to rounding: set it aside instead of failing.
      const bool skip = !done && dep && !has1;
      if (skip) {
        if (sp < 1e3 * kc.ntol_z) {  // not a rounding artefact: give up loudly
          status = QPB_BAD_INPUT;
          done = true;
        }
        ignore |= 1u << pp;
        p = -1;
      }
      const bool full = !done && !dep && (!has1 || t2 <= t1);
      const bool drop = !done && !full && !skip;
      const double t = full ? t2 : (drop ? t1 : 0.0);
      // (5) step
      x = fma(dep ? 0.0 : t, acc, x);
      u = fma(-t, mvN, u);
      up += t;
      // (6) rank-1 updates:  full step: M -= (M n~) z~^T / zeta ;  partial step: M -/+= (..) nu^T / |nu|^2
      double coefP = full ? mvP * izeta : 0.0;
      double coefN = full ? mvN * izeta : 0.0;
      const uint32_t fb = (__ballot_sync(FULL, isP && cons < 0) >> hb) & 0xffffu;  // free slots of this half
      if (full) {
        const int ql = __ffs(fb) - 1;
        if (l == ql) { coefN = -izeta; cons = p; u = up; }
        active |= 1u << p;
        p = -1;
      }
      if (__any_sync(FULL, drop)) {
        if (drop && l == kl) sts12(hs.bv, Mn);
        __syncwarp();
        lds12(drop ? hs.bv : hs.bz, zt);  // dropping halves: zt <- nu = row kl of N~*; others keep z~
        const double gam = dot12(Mn, zt);
        const double idelta = rcp_fast(shfl16(gam, kl & 15));
        const int cdrop = __shfl_sync(FULL, cons, kl & 15, 16);
        if (drop) {
          coefP = isP ? -hs.bv[l] * idelta : 0.0;
          coefN = (cons >= 0) ? gam * idelta : 0.0;
          if (l == kl) { coefN = 1.0; cons = -1; u = 0.0; }
          active &= ~(1u << cdrop);
        }
      }
#pragma unroll
      for (int j = 0; j < 12; j++) {
        Mp[j] = fma(-coefP, zt[j], Mp[j]);
        Mn[j] = fma(-coefN, zt[j], Mn[j]);
      }
    }

    // ---- epilogue: body-frame GRF (:218-232) and tau = J^T f (kinematics.cpp:162-188, 218-231) -----
    const bool good = (status == QPB_OK);
    const double fw = (good && stance) ? x : 0.0;
    const double f0 = shfl16(fw, 3 * leg), f1 = shfl16(fw, 3 * leg + 1), f2 = shfl16(fw, 3 * leg + 2);
    double fb = -1.0 * (hs.rec[ax] * f0 + hs.rec[3 + ax] * f1 + hs.rec[6 + ax] * f2);
    if (!(good && stance)) fb = 0.0;
    const double qa = hs.rec[48 + 3 * leg + ax] + ((ax == 2) ? hs.rec[48 + 3 * leg + 1] : 0.0);
    double sn, cs;
    sincos(ok ? qa : 0.0, &sn, &cs);
    const double s1 = shfl16(sn, 3 * leg), c1 = shfl16(cs, 3 * leg);
    const double s2 = shfl16(sn, 3 * leg + 1), c2 = shfl16(cs, 3 * leg + 1);
    const double s23 = shfl16(sn, 3 * leg + 2), c23 = shfl16(cs, 3 * leg + 2);
    const double fbx = shfl16(fb, 3 * leg), fby = shfl16(fb, 3 * leg + 1), fbz = shfl16(fb, 3 * leg + 2);
    const double l1 = P.link[3 * leg], l2 = P.link[3 * leg + 1], l3 = P.link[3 * leg + 2];
    double Jx, Jy, Jz;
    leg_jacobian_col(ax, l1, l2, l3, s1, c1, s2, c2, s23, c23, Jx, Jy, Jz);
    double tau = Jx * fbx + Jy * fby + Jz * fbz;
    if (P.clamp_tau) tau = fmin(fmax(tau, P.tau_min), P.tau_max);
    if (!(good && stance)) tau = 0.0;
    if (have) store_rec(io, rec, l, fb, tau, status, iters, active | 0x80000000u);
    pair = nwarps + __shfl_sync(FULL, next_ticket, 0);
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

}  // namespace qpb
