// qpb_tpq.cuh -- balance_qp_tpq_kernel: ONE THREAD PER QP, the default kernel when W = w I (the reference's
// configuration) and fzmin >= 0.  The arithmetic is qpb_tpq_core.h (range-space Goldfarb-Idnani, 6x6 Cholesky per
// working-set change); this file is the warp-level plumbing that keeps all 32 lanes busy although iteration counts
// differ from QP to QP (0..40):
//
//   * set-up (load, PD target, lever arms, unconstrained / hinted minimiser) always runs on a FULL warp, 32 new
//     records at a time, and parks the 37-double solver states in a per-warp shared-memory stack ("prep");
//   * the iteration loop runs on whatever the lanes hold; as soon as QPB_TPQ_REFILL lanes are idle, finished lanes
//     push their final working sets (20 doubles) onto a second per-warp stack ("ret") and idle lanes pop fresh states;
//   * the epilogue (polish = one more 6x6 solve on the final faces, world->body, 12 sincos, J^T f, 256-B store) runs
//     on a FULL warp whenever 32 results are parked.
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
#ifndef QPB_TPQ_WARPS
#define QPB_TPQ_WARPS 2
#endif
#ifndef QPB_TPQ_MIN_CTAS
#define QPB_TPQ_MIN_CTAS 4
#endif

namespace qpb {
namespace tpq {

constexpr int PREP_STRIDE = 43;  // doubles per parked solver state (odd: lane-strided access is conflict-free)
constexpr int RET_STRIDE = 21;   // doubles per parked result
constexpr int SIDE_STRIDE = 7;   // per-lane side storage: the right-hand side b (6) of the QP the lane is iterating on
constexpr int PREP_CAP = 32, RET_CAP = 64;

struct __align__(16) WarpStacks {
  double prep[PREP_CAP * PREP_STRIDE];
  double ret[RET_CAP * RET_STRIDE];
  double side[32 * SIDE_STRIDE];
};

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
__device__ __forceinline__ void tpq_load_Rq(const PackedIO& io, int64_t rec, double (&R)[9], double (&q)[12]) {
  const double* p = reinterpret_cast<const double*>(io.in + rec);
#pragma unroll
  for (int j = 0; j < 9; j++) R[j] = __ldg(p + j);
  const double2* pq = reinterpret_cast<const double2*>(p + kQ);
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const double2 t = __ldg(pq + j);
    q[2 * j] = t.x;
    q[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void tpq_load_Rq(const SplitIO& io, int64_t rec, double (&R)[9], double (&q)[12]) {
#pragma unroll
  for (int j = 0; j < 9; j++) R[j] = __ldg(io.Rwb + rec * 9 + j);
#pragma unroll
  for (int j = 0; j < 12; j++) q[j] = __ldg(io.q + rec * 12 + j);
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

// ---- parked solver state: f, r, u, b, then (lo: working set | stance << 24 | status << 28, hi: record index) -----
__device__ __forceinline__ void park_state(double* e, const State& st, const double (&b6)[6], uint32_t rec) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    e[i] = st.f[i];
    e[12 + i] = st.r[i];
    e[24 + i] = st.u[i];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) e[36 + i] = b6[i];
  const uint32_t lo = wset_encode(st.sg) | (st.stance << 24) | ((uint32_t)st.status << 28);
  e[42] = __hiloint2double((int)rec, (int)lo);
}
__device__ __forceinline__ void unpark_state(const double* e, State& st, double* side, uint32_t& rec) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    st.f[i] = e[i];
    st.r[i] = e[12 + i];
    st.u[i] = e[24 + i];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) side[i] = e[36 + i];
  const double w = e[42];
  const uint32_t lo = (uint32_t)__double2loint(w);
  rec = (uint32_t)__double2hiint(w);
  st.stance = (lo >> 24) & 15u;
  st.status = (int)(lo >> 28);
  wset_decode(lo, st.stance, st.sg);
  st.p = -1;
  st.ps = 0.0;
  st.up = 0.0;
  st.iters = 0;
  st.done = st.status != QPB_OK;
}
// ---- parked result: r, b, (lo: working set | stance << 24 | status << 28, hi: record index), iterations ---------------
__device__ __forceinline__ void park_result(double* e, const State& st, const double* side, uint32_t rec) {
#pragma unroll
  for (int i = 0; i < 12; i++) e[i] = st.r[i];
#pragma unroll
  for (int i = 0; i < 6; i++) e[12 + i] = side[i];
  const uint32_t lo = wset_encode(st.sg) | (st.stance << 24) | ((uint32_t)st.status << 28);
  e[18] = __hiloint2double((int)rec, (int)lo);
  e[19] = __hiloint2double(0, st.iters);
}

// Full-warp epilogue over the first cnt parked results.
template <class IO>
__device__ __forceinline__ void flush_results(const qpb_params& P, const FastParams& K, const IO& io, const double* ret, int cnt,
                                              int lane) {
  if (lane < cnt) {
    const double* e = ret + lane * RET_STRIDE;
    State st;
    double b6[6];
#pragma unroll
    for (int i = 0; i < 12; i++) st.r[i] = e[i];
#pragma unroll
    for (int i = 0; i < 6; i++) b6[i] = e[12 + i];
    const double w = e[18];
    const uint32_t rec = (uint32_t)__double2hiint(w);
    const uint32_t lo = (uint32_t)__double2loint(w);
    st.stance = (lo >> 24) & 15u;
    st.status = (int)(lo >> 28);
    st.iters = __double2loint(e[19]);
    wset_decode(lo, st.stance, st.sg);
#pragma unroll
    for (int i = 0; i < 12; i++) st.f[i] = st.u[i] = 0.0;
    polish(K, st, b6);  // the minimiser on the final faces, from scratch
    double R[9], q[12], grf[12], tau[12];
    tpq_load_Rq(io, (int64_t)rec, R, q);
    bool qfin = true;  // the set-up checked slots 0..47; the joint angles are first touched here
#pragma unroll
    for (int i = 0; i < 12; i++) qfin = qfin && (fabs(q[i]) <= 1.79769313486231570e308);
    if (!qfin) st.status = QPB_BAD_INPUT;
    finish(P, R, q, st, grf, tau);
    tpq_store(io, (int64_t)rec, grf, tau, st.status, st.iters, (lo & 0xffffffu) | 0x80000000u);
  }
}

template <class IO>
__global__ void __launch_bounds__(QPB_TPQ_WARPS * 32, QPB_TPQ_MIN_CTAS)
balance_qp_tpq_kernel(const __grid_constant__ qpb_params P, const __grid_constant__ FastParams K, IO io, int64_t n,
                      unsigned long long* __restrict__ ticket) {
  __shared__ WarpStacks stacks[QPB_TPQ_WARPS];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  WarpStacks& ws = stacks[wib];
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t nwarps = gridDim.x * QPB_TPQ_WARPS;
  const uint32_t nchunks = (uint32_t)((n + 31) >> 5);

  uint32_t chunk = blockIdx.x * QPB_TPQ_WARPS + wib;  // next chunk of 32 records this warp sets up
  int prep_n = 0, ret_n = 0;                           // warp-uniform stack heights
  bool have = false;                                   // this lane holds a QP
  uint32_t rec = 0;
  State st;
  st.done = true;
  st.status = QPB_OK;
  st.iters = 0;
  st.p = -1;
  st.ps = st.up = 0.0;
  st.stance = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) st.f[i] = st.u[i] = st.sg[i] = st.r[i] = 0.0;

  for (;;) {
    const uint32_t busy = __ballot_sync(FULL, have && !st.done);
    if (32 - __popc(busy) >= QPB_TPQ_REFILL || busy == 0u) {
      // ---- retire: finished lanes park their results; 32 parked results -> full-warp epilogue ----------------
      const bool fin = have && st.done;
      const uint32_t fm = __ballot_sync(FULL, fin);
      if (fm) {
        if (fin) {
          park_result(ws.ret + (ret_n + __popc(fm & lt)) * RET_STRIDE, st, ws.side + lane * SIDE_STRIDE, rec);
          have = false;
        }
        ret_n += __popc(fm);
        __syncwarp();
        if (ret_n >= 32) {
          flush_results(P, K, io, ws.ret, 32, lane);
          __syncwarp();
          ret_n -= 32;
          if (lane < ret_n) {  // move the remainder down
            double t[20];
#pragma unroll
            for (int i = 0; i < 20; i++) t[i] = ws.ret[(32 + lane) * RET_STRIDE + i];
#pragma unroll
            for (int i = 0; i < 20; i++) ws.ret[lane * RET_STRIDE + i] = t[i];
          }
          __syncwarp();
        }
      }
      // ---- refill: idle lanes pop parked states; an empty stack is restocked by a full-warp set-up --------------
      bool stocked = false;
#pragma unroll 1
      for (int pass = 0; pass < 2; pass++) {
        const uint32_t idle = ~__ballot_sync(FULL, have);
        const int want = __popc(idle);
        if (want == 0) break;
        if (prep_n == 0) {
          if (stocked || chunk >= nchunks) break;
          stocked = true;
          const int64_t r0 = (int64_t)chunk * 32 + lane;
          const bool valid = r0 < n;
          uint32_t nt = 0;
          if (lane == 0) nt = (uint32_t)atomicAdd(ticket, 1ULL);
          if (valid) {
            double v[48];
            uint32_t cbytes, hint;
            tpq_load(io, r0, v, cbytes, hint);
            State s0;
            double b6[6];
            setup(P, K, v, cbytes, hint, s0, b6);
            park_state(ws.prep + lane * PREP_STRIDE, s0, b6, (uint32_t)r0);
          }
          prep_n = __popc(__ballot_sync(FULL, valid));  // valid lanes are the low ones: the stack is dense
          chunk = nwarps + __shfl_sync(FULL, nt, 0);
          __syncwarp();
        }
        const int rank = __popc(idle & lt);
        if (!have && rank < prep_n) {
          unpark_state(ws.prep + (prep_n - 1 - rank) * PREP_STRIDE, st, ws.side + lane * SIDE_STRIDE, rec);
          have = true;
        }
        prep_n -= min(want, prep_n);
        __syncwarp();
      }
      if (__ballot_sync(FULL, have) == 0u) {  // nothing left anywhere
        if (ret_n > 0) flush_results(P, K, io, ws.ret, ret_n, lane);
        break;
      }
    }
    iterate(K, st);
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
