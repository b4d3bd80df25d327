// qpb_tpq.cuh -- balance_qp_tpq_kernel: the default balance kernel when W = w I (the reference's configuration) and
// fzmin >= 0.  The arithmetic is qpb_tpq_core.h (range-space Goldfarb-Idnani, a 6x6 Cholesky per working-set change).
// A QP is iterated on by LPQ lanes (template parameter: 1 = one thread per QP, 2 = two legs per lane, 4 = one leg per
// lane); this file is the warp-level plumbing that keeps the lanes busy although iteration counts differ from QP to QP
// (0..40):
//
//   * set-up (load, PD target, lever arms, unconstrained / hinted minimiser) always runs one record per THREAD on a full
//     warp, 32 new records at a time, and parks the 64-double solver states in a per-warp shared-memory stack ("prep");
//   * the iteration loop runs on the 32 / LPQ QPs the warp holds; as soon as QPB_TPQ_REFILL lanes are idle, finished QPs
//     push their final working sets (8 doubles) onto a second per-warp stack ("ret") and idle lanes pop fresh states;
//   * the epilogue (polish = one more 6x6 solve on the final faces, world->body, 12 sincos, J^T f, 256-B store) runs
//     one result per thread on a full warp whenever 32 results are parked.
//
// Warps never synchronise with each other: the stacks are private to a warp (__syncwarp only), work is claimed in
// chunks of 32 records from a global ticket.
#pragma once

#include <cuda_runtime.h>

#include "qpb_kernel.cuh"
#include "qpb_tpq_core.h"

#ifndef QPB_TPQ_REFILL
#define QPB_TPQ_REFILL 2  // idle lanes that trigger a retire + refill
#endif

namespace qpb {
namespace tpq {

constexpr int PREP_STRIDE = 65;  // doubles per parked solver state (odd: lane-strided access is conflict-free)
constexpr int RET_STRIDE = 9;    // doubles per parked result
constexpr int SIDE_STRIDE = kSideSize;  // per-QP storage while it is iterated on: b (6), G (21), lever arms (12)
constexpr int PREP_CAP = 32, RET_CAP = 32;

// Shared memory of a CTA of W warps: ONE stack of prepared states for all its warps (popped and restocked under a
// try-lock: a warp that finds it taken just keeps iterating and tries again a round later), and per warp the stack of
// results awaiting the epilogue and the b / G blocks of the QPs it is iterating on.
template <int LPQ, int W>
struct __align__(16) CtaShared {
  double prep[PREP_CAP * PREP_STRIDE];
  struct PerWarp {
    double ret[RET_CAP * RET_STRIDE];
    double side[(32 / LPQ) * SIDE_STRIDE];
  } w[W];
  int lock;       // 0 free, 1 held
  int prep_n;     // height of prep (read and written under the lock)
  int exhausted;  // the work ticket has run past the last chunk (set under the lock, never cleared)
};

// kernel shape per lanes-per-QP: warps per CTA and the minimum CTAs per SM (= the register cap)
template <int LPQ> struct Shape;
template <> struct Shape<1> { static constexpr int W = 2, MIN_CTAS = 4; };   // 255 registers,  8 warps / SM
template <> struct Shape<2> { static constexpr int W = 4, MIN_CTAS = 3; };   // 168 registers, 12 warps / SM
template <> struct Shape<4> { static constexpr int W = 4, MIN_CTAS = 4; };   // 128 registers, 16 warps / SM

// ---- exchanges between the LPQ lanes of a QP (xor butterflies inside aligned groups of LPQ lanes) -----------------------
template <int LPQ>
__device__ __forceinline__ uint32_t group_umax(uint32_t v) {
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
template <int LPQ>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// smallest fraction ub / rb over the group (rb = 0: none); ties go to the smaller row index so all lanes agree
template <int LPQ>
__device__ __forceinline__ void group_min_ratio(double& ub, double& rb, int& kb) {
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) {
    const double ub2 = __shfl_xor_sync(FULL, ub, o), rb2 = __shfl_xor_sync(FULL, rb, o);
    const int kb2 = __shfl_xor_sync(FULL, kb, o);
    const double lhs = ub2 * rb, rhs = ub * rb2;
    if (hi32(rb2) > 0 && (lhs < rhs || (lhs == rhs && hi32(rb) > 0 && kb2 < kb))) {
      ub = ub2;
      rb = rb2;
      kb = kb2;
    }
  }
}

// One working-set change for every QP the warp holds (or the optimality test that ends a solve).
template <int LPQ>
__device__ __forceinline__ void iterate_group(const FastParams& K, Lane<4 / LPQ>& ln, int j, double* side) {
  constexpr int LPL = 4 / LPQ;
  const uint32_t best = group_umax<LPQ>(select_local<LPL>(K, ln, j));
  bool fresh;
  const double slack = group_sum<LPQ>(select_commit<LPL>(K, ln, j, best, fresh));
  if (fresh) ln.sp = slack;
  StepTmp<LPL> T;
  double ub, rb;
  int kb;
  direction<LPL>(K, ln, j, side, T, ub, rb, kb);
  group_min_ratio<LPQ>(ub, rb, kb);
  advance<LPL>(K, ln, j, side, T, ub, rb, kb, j == 0);
}

// ---- record access: 48 doubles + contact bytes + warm-start word in, 256-B record out --------------------------
__device__ __forceinline__ void tpq_load(const PackedIO& io, int64_t rec, double (&v)[48], uint32_t& cbytes, uint32_t& hint) {
  const double2* p = reinterpret_cast<const double2*>(io.in + rec);
#pragma unroll
  for (int j = 0; j < 24; j++) {
    const double2 t = __ldg(p + j);
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
  const uint2 c = __ldg(reinterpret_cast<const uint2*>(io.in + rec) + 60);  // contact[4], pad[0..3] = warm-start word
  cbytes = c.x;
  hint = c.y;
}
__device__ __forceinline__ void tpq_load(const SplitIO& io, int64_t rec, double (&v)[48], uint32_t& cbytes, uint32_t& hint) {
#pragma unroll
  for (int m = 0; m < 48; m++) v[m] = split_slot(io, rec, m);
  cbytes = io.contact[rec * 4] | (io.contact[rec * 4 + 1] << 8) | (io.contact[rec * 4 + 2] << 16) |
           ((uint32_t)io.contact[rec * 4 + 3] << 24);
  hint = 0u;
}
// what the epilogue needs of a record: R (9), feet (12), q (12)
__device__ __forceinline__ void tpq_load_Rq(const PackedIO& io, int64_t rec, double (&R)[9], double (&feet)[12], double (&q)[12]) {
  const double* p = reinterpret_cast<const double*>(io.in + rec);
#pragma unroll
  for (int j = 0; j < 9; j++) R[j] = __ldg(p + j);
  const double2* pq = reinterpret_cast<const double2*>(p + kFeet);  // feet, q: slots 36..59
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const double2 t = __ldg(pq + j), t2 = __ldg(pq + 6 + j);
    feet[2 * j] = t.x;
    feet[2 * j + 1] = t.y;
    q[2 * j] = t2.x;
    q[2 * j + 1] = t2.y;
  }
}
__device__ __forceinline__ void tpq_load_Rq(const SplitIO& io, int64_t rec, double (&R)[9], double (&feet)[12], double (&q)[12]) {
#pragma unroll
  for (int j = 0; j < 9; j++) R[j] = __ldg(io.Rwb + rec * 9 + j);
#pragma unroll
  for (int j = 0; j < 12; j++) {
    feet[j] = __ldg(io.feet + rec * 12 + j);
    q[j] = __ldg(io.q + rec * 12 + j);
  }
}
__device__ __forceinline__ void tpq_store(const PackedIO& io, int64_t rec, const double (&grf)[12], const double (&tau)[12],
                                          int status, int iters, uint32_t wword) {
  double2* o = reinterpret_cast<double2*>(io.out + rec);
#pragma unroll
  for (int j = 0; j < 6; j++) o[j] = make_double2(grf[2 * j], grf[2 * j + 1]);
#pragma unroll
  for (int j = 0; j < 6; j++) o[6 + j] = make_double2(tau[2 * j], tau[2 * j + 1]);
  int4* t = reinterpret_cast<int4*>(o + 12);  // bytes 192..255: status, iters, working-set word, zero padding
  t[0] = make_int4(status, iters, (int)wword, 0);
  t[1] = make_int4(0, 0, 0, 0);
  t[2] = make_int4(0, 0, 0, 0);
  t[3] = make_int4(0, 0, 0, 0);
}
__device__ __forceinline__ void tpq_store(const SplitIO& io, int64_t rec, const double (&grf)[12], const double (&tau)[12],
                                          int status, int /*iters*/, uint32_t /*wword*/) {
#pragma unroll
  for (int j = 0; j < 12; j++) io.grf[rec * 12 + j] = grf[j];
  if (io.tau) {
#pragma unroll
    for (int j = 0; j < 12; j++) io.tau[rec * 12 + j] = tau[j];
  }
  if (io.status) io.status[rec] = status;
}

// ---- parked solver state: f, r, u, b, G, then (lo: working set | stance << 24 | status << 28, hi: record index) --
__device__ __forceinline__ void park_state(double* e, const State& st, const double (&b6)[6], const double (&G)[21], uint32_t rec) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    e[i] = st.f[i];
    e[12 + i] = st.r[i];
    e[24 + i] = st.u[i];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) e[36 + i] = b6[i];
#pragma unroll
  for (int i = 0; i < 21; i++) e[42 + i] = G[i];
  const uint32_t lo = st.word | (st.stance << 24) | ((uint32_t)st.status << 28);
  e[63] = __hiloint2double((int)rec, (int)lo);
}
// a group of LPQ lanes takes a parked state: every lane its share of f and u, all of r; b and G go to the QP's side block
template <int LPQ>
__device__ __forceinline__ void unpark_state(const double* e, Lane<4 / LPQ>& ln, int j, double* side, uint32_t& rec) {
  const double w = e[63];
  const uint32_t lo = (uint32_t)__double2loint(w);
  rec = (uint32_t)__double2hiint(w);
  lane_init<4 / LPQ>(ln, j, e, e + 12, e + 24, lo & 0xffffffu, (lo >> 24) & 15u, (int)(lo >> 28));
#pragma unroll
  for (int i = 0; i < (27 + LPQ - 1) / LPQ; i++) {
    const int k = j + LPQ * i;
    if (k < 27) side[k] = e[36 + k];  // b, G
  }
#pragma unroll
  for (int i = 0; i < 12 / LPQ; i++) side[kSideR + j + LPQ * i] = e[12 + j + LPQ * i];  // all lever arms
}
// ---- parked result: b, (lo: working set | stance << 24 | status << 28, hi: record index), iterations ------------------
template <int LPL>
__device__ __forceinline__ void park_result(double* e, const Lane<LPL>& ln, const double* side, uint32_t rec) {
#pragma unroll
  for (int i = 0; i < 6; i++) e[i] = side[i];
  const uint32_t lo = ln.word | (ln.stance << 24) | ((uint32_t)ln.status << 28);
  e[6] = __hiloint2double((int)rec, (int)lo);
  e[7] = __hiloint2double(0, ln.iters);
}

// Full-warp epilogue over the first cnt parked results: polish (the minimiser on the final faces, from scratch),
// world->body, J^T f, store.
template <class IO>
__device__ __noinline__ void flush_results(const qpb_params& P, const FastParams& K, const IO& io, const double* ret, int cnt,
                                           int lane) {
  if (lane < cnt) {
    const double* e = ret + lane * RET_STRIDE;
    State st;
    double b6[6];
#pragma unroll
    for (int i = 0; i < 6; i++) b6[i] = e[i];
    const double w = e[6];
    const uint32_t rec = (uint32_t)__double2hiint(w);
    const uint32_t lo = (uint32_t)__double2loint(w);
    st.word = lo & 0xffffffu;
    st.stance = (lo >> 24) & 15u;
    st.status = (int)(lo >> 28);
    st.iters = __double2loint(e[7]);
    double R[9], feet[12], q[12], grf[12], tau[12];
    tpq_load_Rq(io, (int64_t)rec, R, feet, q);
    bool qfin = true;  // the set-up checked slots 0..47; the joint angles are first touched here
#pragma unroll
    for (int i = 0; i < 12; i++) qfin = qfin && (fabs(q[i]) <= 1.79769313486231570e308);
    if (!qfin && st.status == QPB_OK) st.status = QPB_BAD_INPUT;
    const bool sane = st.status != QPB_BAD_INPUT;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int k = 0; k < 3; k++)  // lever arms r_i = R p_i again (cheaper than parking them)
        st.r[3 * i + k] = sane ? R[3 * k] * feet[3 * i] + R[3 * k + 1] * feet[3 * i + 1] + R[3 * k + 2] * feet[3 * i + 2] : 0.0;
#pragma unroll
    for (int i = 0; i < 12; i++) st.f[i] = st.u[i] = 0.0;
    polish(K, st, b6);
    finish(P, R, q, st, grf, tau);
    tpq_store(io, (int64_t)rec, grf, tau, st.status, st.iters, st.word | 0x80000000u);
  }
}

template <class IO, int LPQ>
__global__ void __launch_bounds__(Shape<LPQ>::W * 32, Shape<LPQ>::MIN_CTAS)
balance_qp_tpq_kernel(const __grid_constant__ qpb_params P, const __grid_constant__ FastParams K, IO io, int64_t n,
                      unsigned long long* __restrict__ ticket) {
  constexpr int LPL = 4 / LPQ, W = Shape<LPQ>::W;
  __shared__ CtaShared<LPQ, W> sh;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sh.lock = 0;
    sh.prep_n = 0;
    sh.exhausted = 0;
  }
  __syncthreads();
  typename CtaShared<LPQ, W>::PerWarp& ws = sh.w[wib];
  const int j = lane & (LPQ - 1);                             // lane within its QP
  const uint32_t leaders = 0xffffffffu / ((1u << LPQ) - 1u);  // bit of the first lane of every group
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t glt = (1u << (lane & ~(LPQ - 1))) - 1u;  // lanes below this lane's group
  double* side = ws.side + (lane / LPQ) * SIDE_STRIDE;
  const uint32_t nchunks = (uint32_t)((n + 31) >> 5);
  constexpr int kRefill = (QPB_TPQ_REFILL + LPQ - 1) / LPQ;  // idle QP slots that trigger a retire + refill

  int ret_n = 0;      // warp-uniform height of the result stack
  bool have = false;  // this lane's group holds a QP
  uint32_t rec = 0;
  Lane<LPL> ln;
  {
    const double zero[12] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
    lane_init<LPL>(ln, j, zero, zero, zero, 0u, 0u, QPB_OK);
    ln.done = true;
  }

  for (;;) {
    const uint32_t busy = __ballot_sync(FULL, have && !ln.done) & leaders;
    if (32 / LPQ - __popc(busy) >= kRefill || busy == 0u) {
      // ---- retire: finished QPs park their results; a full stack -> full-warp epilogue ---------------------------
      const bool fin = have && ln.done;
      const uint32_t fm = __ballot_sync(FULL, fin) & leaders;
      if (fm) {
        if (ret_n + __popc(fm) > RET_CAP) {  // no room for all of them: run the epilogue on what is parked (>= 31/32 full)
          flush_results(P, K, io, ws.ret, ret_n, lane);
          ret_n = 0;
          __syncwarp();
        }
        if (fin) {
          if (j == 0) park_result<LPL>(ws.ret + (ret_n + __popc(fm & lt)) * RET_STRIDE, ln, side, rec);
          have = false;
        }
        ret_n += __popc(fm);
        __syncwarp();
        if (ret_n == RET_CAP) {
          flush_results(P, K, io, ws.ret, ret_n, lane);
          ret_n = 0;
          __syncwarp();
        }
      }
      // ---- refill under the CTA's try-lock: idle groups pop prepared states; an empty stack is restocked by a set-up
      //      of the next 32 records, one record per thread ----------------------------------------------------------------
      int got = 0;
      if (lane == 0) got = atomicCAS(&sh.lock, 0, 1) == 0;
      got = __shfl_sync(FULL, got, 0);
      if (got) {
        __threadfence_block();
        int prep_n = *(volatile int*)&sh.prep_n;
        int exhausted = *(volatile int*)&sh.exhausted;
        bool stocked = false;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
          const uint32_t idle = ~__ballot_sync(FULL, have) & leaders;
          const int want = __popc(idle);
          if (want == 0) break;
          if (prep_n == 0) {
            if (stocked || exhausted) break;
            stocked = true;
            uint32_t chunk = 0;
            if (lane == 0) chunk = (uint32_t)atomicAdd(ticket, 1ULL);
            chunk = __shfl_sync(FULL, chunk, 0);
            if (chunk >= nchunks) {
              exhausted = 1;
              break;
            }
            const int64_t r0 = (int64_t)chunk * 32 + lane;
            const bool valid = r0 < n;
            if (valid) {
              double v[48];
              uint32_t cbytes, hint;
              tpq_load(io, r0, v, cbytes, hint);
              State s0;
              double b6[6], G[21];
              setup(P, K, v, cbytes, hint, s0, b6, G);
              park_state(sh.prep + lane * PREP_STRIDE, s0, b6, G, (uint32_t)r0);
            }
            prep_n = __popc(__ballot_sync(FULL, valid));  // valid lanes are the low ones: the stack is dense
            __syncwarp();
          }
          const int rank = __popc(idle & glt);
          if (!have && rank < prep_n) {
            unpark_state<LPQ>(sh.prep + (prep_n - 1 - rank) * PREP_STRIDE, ln, j, side, rec);
            have = true;
          }
          prep_n -= min(want, prep_n);
          __syncwarp();
        }
        if (lane == 0) {
          *(volatile int*)&sh.prep_n = prep_n;
          *(volatile int*)&sh.exhausted = exhausted;
          __threadfence_block();
          atomicExch(&sh.lock, 0);
        }
        __syncwarp();
      }
      if (__ballot_sync(FULL, have) == 0u) {
        // This warp holds nothing.  It is finished once the input is exhausted and the shared stack is empty (exhausted
        // is read first: after it is set no set-up runs, so prep_n can only go down).
        const int ex = *(volatile int*)&sh.exhausted;
        const int pn = *(volatile int*)&sh.prep_n;
        if (ex && pn == 0) {
          if (ret_n > 0) flush_results(P, K, io, ws.ret, ret_n, lane);
          break;
        }
        continue;  // the lock was busy or another warp is restocking: look again
      }
    }
    iterate_group<LPQ>(K, ln, j, side);
    __syncwarp();  // G written by the first lane of a QP is read by its other lanes in the next round
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

}  // namespace tpq
}  // namespace qpb
