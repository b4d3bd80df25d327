// qpb_tpq.cuh -- the default balance path when W = w I (the reference's configuration) and fzmin >= 0: three kernels
// around the arithmetic of qpb_tpq_core.h (range-space Goldfarb-Idnani, one 6x6 Cholesky per working-set change).
//
//   tpq_setup_kernel   one THREAD per record: load, PD target, lever arms, the unconstrained (or hinted) minimiser;
//                      writes a 512-B prepared record (f, r, u, b, G, working set) to scratch memory
//   tpq_loop_kernel    the active-set loop.  A QP is iterated on by LPQ lanes (template: 1 = one thread per QP, 2 = two
//                      legs per lane, 4 = one leg per lane).  Persistent warps: whenever QPB_TPQ_REFILL lanes are idle,
//                      finished QPs write their final working set (8 bytes) back and idle lanes claim the next prepared
//                      records from a global ticket -- iteration counts differ from QP to QP (0..40), lanes never wait
//                      for each other beyond that threshold, warps never synchronise with each other
//   tpq_finish_kernel  one THREAD per record: polish (the minimiser on the final faces from one fresh 6x6 solve),
//                      world->body, 12 sincos, J^T f, the 256-B result record
//
// Splitting the path in three lets each part run at its own register budget and keeps the loop's code small enough for
// the instruction cache: fused in one kernel (an earlier build, profiles/r02_ncu_tpq_v3_*) every variant sat at the
// same 4.3e8 QP/s on config 3, latency-bound at 8-16 warps per SM with the set-up's and epilogue's registers.
#pragma once

#include <cuda_runtime.h>

#include "qpb_kernel.cuh"
#include "qpb_tpq_core.h"

#ifndef QPB_TPQ_REFILL
#define QPB_TPQ_REFILL 2  // idle lanes that trigger a retire + refill in the loop kernel
#endif

namespace qpb {
namespace tpq {

constexpr int kEdgeThreads = 128;   // set-up / finishing kernels

// Programmatic dependent launch (sm_90+).  The three passes of a batch are launched back to back on one stream, and between
// two of them the GPU used to sit idle for 4-8 us (per-CTA timelines, profiles/r02_overlap_ab.txt: loop pass over at 71.7 us,
// first CTA of the finishing pass at 75.8-79.9 us) -- 6-10 % of a 65 536-record step.  Now the set-up and loop kernels tell the
// hardware at their very start that their dependents may be SCHEDULED (pdl_release), and the loop and finishing kernels are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and begin with pdl_wait(), which returns once the whole
// preceding grid has completed and its memory operations are visible: their CTAs move onto the SMs as the predecessor's CTAs
// leave and sit in that wait -- issuing nothing, so the long QPs that end the loop pass are not slowed (unlike a finishing pass
// that really runs beside them, same file) -- and start the moment the predecessor is done.  Results cannot change: nothing of a
// predecessor is read before the wait.  Launched without the attribute (stream capture, QPB_TPQ_PDL=0) both are no-ops.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifdef QPB_TPQ_TIMELINE
// developer build only (tools/timeline.py): when every CTA of the loop and finishing passes started and ended (ns, %globaltimer)
__device__ unsigned long long g_tl[4][8192];  // loop start, loop end, finish start, finish end -- by blockIdx.x
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define QPB_TL(k) do { if (threadIdx.x == 0 && blockIdx.x < 8192) g_tl[k][blockIdx.x] = tl_now(); } while (0)
#else
#define QPB_TL(k) do { } while (0)
#endif

// minimum CTAs per SM = the register cap (65536 / (128 threads * MIN_CTAS)); overridable for experiments
#ifndef QPB_TPQ_MINCTAS_1
#define QPB_TPQ_MINCTAS_1 2  // 255 registers
#endif
#ifndef QPB_TPQ_MINCTAS_2
#define QPB_TPQ_MINCTAS_2 3  // 168
#endif
#ifndef QPB_TPQ_MINCTAS_4
#define QPB_TPQ_MINCTAS_4 4  // 128
#endif
#ifndef QPB_TPQ_EDGE_MINCTAS
#define QPB_TPQ_EDGE_MINCTAS 2
#endif
#ifndef QPB_TPQ_FINISH_MINCTAS
#define QPB_TPQ_FINISH_MINCTAS QPB_TPQ_EDGE_MINCTAS
#endif
// loop kernel shape per lanes-per-QP: threads per CTA (static shared memory must stay under 48 KB), minimum CTAs per SM,
// records per staged batch
template <int LPQ> struct LoopShape;
template <> struct LoopShape<1> { static constexpr int THREADS = 64, MIN_CTAS = 2 * QPB_TPQ_MINCTAS_1, STAGE = 16; };
template <> struct LoopShape<2> { static constexpr int THREADS = 128, MIN_CTAS = QPB_TPQ_MINCTAS_2, STAGE = 12; };
template <> struct LoopShape<4> { static constexpr int THREADS = 128, MIN_CTAS = QPB_TPQ_MINCTAS_4, STAGE = 8; };

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

// One working-set change for every QP the warp holds, then the choice of the next row (or the end of the solve: nothing
// is violated any more).  Every QP enters with its row already chosen -- the first one by the set-up pass.
template <int LPQ>
__device__ __forceinline__ void iterate_group(const FastParams& K, Lane<4 / LPQ>& ln, int j, double* side) {
  constexpr int LPL = 4 / LPQ;
  StepTmp<LPL> T;
  double ub, rb;
  int kb;
  direction<LPL>(K, ln, j, side, T, ub, rb, kb);
  group_min_ratio<LPQ>(ub, rb, kb);
  advance<LPL>(K, ln, j, side, T, ub, rb, kb, j == 0);
  const uint32_t best = group_umax<LPQ>(select_local<LPL>(K, ln, j));
  bool fresh;
  const double slack = group_sum<LPQ>(select_commit<LPL>(K, ln, j, best, fresh));
  if (fresh) ln.sp = slack;
}

// meta word of a prepared record.  low: working set (24) | stance (4) | status (2).  high: working-set changes so far
// (16) | first row of the loop (5 bits: 3 leg + group, + 16 for row B) << 16 | "there is one" << 21.
__device__ __forceinline__ double pack_meta(uint32_t word, uint32_t stance, int status, int iters, uint32_t key) {
  const uint32_t lo = word | (stance << 24) | ((uint32_t)status << 28);
  const uint32_t hi = ((uint32_t)iters & 0xffffu) | ((key & 31u) << 16) | ((key >> 31) << 21);
  return __hiloint2double((int)hi, (int)lo);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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
// what the finishing pass needs of a record: R (9), q (12)
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
// start of a packed record as doubles (only meaningful for PackedIO; the other loader never stages)
__device__ __forceinline__ const double* packed_base(const PackedIO& io, int64_t rec) { return reinterpret_cast<const double*>(io.in + rec); }
__device__ __forceinline__ const double* packed_base(const SplitIO&, int64_t) { return nullptr; }
// joint angle i of a record whose slots 0..47 are already in registers (packed records carry q in slots 48..59)
__device__ __forceinline__ double tpq_q(const PackedIO& io, int64_t rec, const double (&)[48], int i) {
  return __ldg(reinterpret_cast<const double*>(io.in + rec) + kQ + i);
}
__device__ __forceinline__ double tpq_q(const SplitIO& io, int64_t rec, const double (&)[48], int i) { return __ldg(io.q + rec * 12 + i); }
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

// finish() rolled over the legs (ONE copy of the sincos code) with its per-leg arrays in a column of shared memory
// instead of local memory: element i of this thread's array X sits at col[(X + i) * stride].  In: forces at F, joint
// angles at Q.  Out: body-frame forces at GRF (may alias F), torques at TAU.  Same arithmetic as finish().
template <int F, int Q, int GRF, int TAU, class Params>
__device__ __forceinline__ void finish_rolled_smem(const Params& P, const double* R, double* col, int stride, int status, uint32_t stance) {
  const bool good = status == QPB_OK, qok = status != QPB_BAD_INPUT;
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    const bool on = good && ((stance >> i) & 1u);
    const double f0 = col[(F + 3 * i) * stride], f1 = col[(F + 3 * i + 1) * stride], f2 = col[(F + 3 * i + 2) * stride];
    double fb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) fb[k] = on ? -1.0 * (R[k] * f0 + R[3 + k] * f1 + R[6 + k] * f2) : 0.0;
    const double q0 = col[(Q + 3 * i) * stride], q1 = col[(Q + 3 * i + 1) * stride], q2 = col[(Q + 3 * i + 2) * stride];
    double s1, c1, s2, c2, s23, c23;
    sincos(qok ? q0 : 0.0, &s1, &c1);
    sincos(qok ? q1 : 0.0, &s2, &c2);
    sincos(qok ? q1 + q2 : 0.0, &s23, &c23);
    double tq[3];
    leg_jt(P.link[3 * i], P.link[3 * i + 1], P.link[3 * i + 2], s1, c1, s2, c2, s23, c23, fb[0], fb[1], fb[2], tq);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double v = tq[k];
      if (P.clamp_tau) v = fmin(fmax(v, P.tau_min), P.tau_max);
      col[(GRF + 3 * i + k) * stride] = fb[k];
      col[(TAU + 3 * i + k) * stride] = on ? v : 0.0;
    }
  }
}

// ---- pass 1: set-up, one thread per record ----------------------------------------------------------------------------
// Writes a dual-feasible starting pair into the prepared record: f, u, G and the meta word (working set, contact mask,
// status, block rounds spent).  Called for every pair the block rounds of start() accept; the last call wins.
struct PrepCommit {
  double* e;
  uint32_t key;  // first row of the loop at the last pair committed (0: that pair is optimal, no loop needed)
  double meta;   // meta word of the last pair committed
  __device__ __forceinline__ void operator()(const State& st, const double (&G)[21], uint32_t k) {
    key = k;
    meta = pack_meta(st.word, st.stance, st.status, st.iters, k);
    if (k == 0u) return;  // optimal (or unsolvable) already: the loop never sees this record, the finishing pass needs only r, b
    double2* o = reinterpret_cast<double2*>(e);
#pragma unroll
    for (int i = 0; i < 6; i++) {
      o[kPrepF / 2 + i] = make_double2(st.f[2 * i], st.f[2 * i + 1]);
      o[kPrepU / 2 + i] = make_double2(st.u[2 * i], st.u[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 10; i++) o[kPrepG / 2 + i] = make_double2(G[2 * i], G[2 * i + 1]);
    o[kPrepG / 2 + 10] = make_double2(G[20], meta);
  }
};

// Staged records (packed records only; QPB_TPQ_STAGE_IN=0 turns it off): the 32 records of a warp are copied to shared
// memory by 32 cp.async instructions, each moving one whole record (31 lanes x 16 B, coalesced), instead of 25 loads per
// thread that each touch 32 different lines; the threads read their record with conflict-free 128-bit shared loads.
// The same row then collects the prepared record -- every pair the block rounds commit lands in shared memory instead of
// global memory -- and leaves in one coalesced copy per record.  Config 3 +10 %, config 2 +3 % (input alone: +2 % / 0;
// profiles/r02_setup_staging_ab.txt): `lg_throttle` and most of `long_scoreboard` of the set-up pass were these accesses.
#ifndef QPB_TPQ_STAGE_IN
#define QPB_TPQ_STAGE_IN 1
#endif
constexpr int kStageStride = 66;  // doubles per staged record (62 used; 132 words: a quarter-warp of LDS.128 covers all banks)
template <class IO> struct StageIn { static constexpr bool on = false; };
#if QPB_TPQ_STAGE_IN
template <> struct StageIn<PackedIO> { static constexpr bool on = true; };
#endif
constexpr size_t kSetupStageBytes = (size_t)kEdgeThreads * kStageStride * sizeof(double);

// EARLY (warm batches): a record whose starting pair is already optimal -- nearly all of them when the records carry last
// tick's working sets -- is finished right here (its forces ARE the minimiser on the final faces from one fresh solve):
// epilogue, result record, no scratch traffic; only the others get a prepared record and a place in the worklist, and
// the finishing pass then walks the worklist instead of the batch (tpq_finish_kernel<IO, true>).
template <class IO, bool EARLY>
__global__ void __launch_bounds__(kEdgeThreads, QPB_TPQ_EDGE_MINCTAS)
tpq_setup_kernel(const __grid_constant__ EdgeParams P, const __grid_constant__ FastParams K, IO io, int64_t n, double* __restrict__ prep,
                 double* __restrict__ res, uint32_t* __restrict__ work, unsigned long long* __restrict__ ticket) {
  pdl_release();
  const int64_t rec = (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x;
  bool need = false;  // this record goes through the active-set loop (its starting pair is not optimal yet)
  extern __shared__ __align__(16) double stage_dyn[];
  const double* mine = nullptr;  // this thread's staged record
  if (StageIn<IO>::on) {
    const int lane = threadIdx.x & 31;
    double* stage = stage_dyn + (threadIdx.x >> 5) * 32 * kStageStride;
    const int64_t base = (int64_t)blockIdx.x * kEdgeThreads + (threadIdx.x & ~31);
    const int cnt = n - base >= 32 ? 32 : (n > base ? (int)(n - base) : 0);
    if (lane < 31)
      for (int e = 0; e < cnt; e++) cp_async16(stage + e * kStageStride + 2 * lane, packed_base(io, base + e) + 2 * lane);
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    mine = stage + lane * kStageStride;
  }
  if (rec < n) {
    double v[48];
    uint32_t cbytes, hint;
    if (StageIn<IO>::on) {
#pragma unroll
      for (int j = 0; j < 24; j++) {
        const double2 t = reinterpret_cast<const double2*>(mine)[j];
        v[2 * j] = t.x;
        v[2 * j + 1] = t.y;
      }
      const uint2 c = reinterpret_cast<const uint2*>(mine)[60];
      cbytes = c.x;
      hint = c.y;
    } else {
      tpq_load(io, rec, v, cbytes, hint);
    }
    State st;
    double b6[6], G[21];
    __shared__ double cols[EARLY ? kEdgeThreads * 36 : 1];  // per thread: f -> grf (12), q (12), tau (12)
    double* col = cols + threadIdx.x;
    bool qfin = true;
    if (EARLY && StageIn<IO>::on) {  // the staged row becomes the prepared record: take the joint angles out first
#pragma unroll
      for (int i = 0; i < 12; i++) {
        const double qi = mine[kQ + i];
        qfin = qfin && (fabs(qi) <= 1.79769313486231570e308);
        col[(12 + i) * kEdgeThreads] = qi;
      }
    }
    // staged: the prepared record is put together in this thread's row of shared memory (every pair the block rounds
    // commit lands there, not in global memory) and leaves in one coalesced copy per record below
    PrepCommit commit{ StageIn<IO>::on ? const_cast<double*>(mine) : prep + rec * kPrepSize, 0u, 0.0 };
    setup(P, K, v, cbytes, hint, st, b6, G, commit);
    need = commit.key != 0u;
    if (EARLY && !need) {
      // (the pair committed last is the one st holds: start() stops at the first optimal pair)
#pragma unroll
      for (int i = 0; i < 12; i++) {
        if (!StageIn<IO>::on) {
          const double qi = tpq_q(io, rec, v, i);
          qfin = qfin && (fabs(qi) <= 1.79769313486231570e308);
          col[(12 + i) * kEdgeThreads] = qi;
        }
        col[i * kEdgeThreads] = st.f[i];
      }
      if (!qfin && st.status == QPB_OK) st.status = QPB_BAD_INPUT;
      finish_rolled_smem<0, 12, 0, 24>(P, v + kR, col, kEdgeThreads, st.status, st.stance);
      double grf[12], tau[12];
#pragma unroll
      for (int i = 0; i < 12; i++) {
        grf[i] = col[i * kEdgeThreads];
        tau[i] = col[(24 + i) * kEdgeThreads];
      }
      tpq_store(io, rec, grf, tau, st.status, st.iters, st.word | 0x80000000u);
    } else {
      // what does not depend on the working set: lever arms and the right-hand side (after a non-finite input they are zero)
      double2* o = reinterpret_cast<double2*>(commit.e);
#pragma unroll
      for (int i = 0; i < 6; i++) o[kPrepR / 2 + i] = make_double2(st.r[2 * i], st.r[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 3; i++) o[kPrepB / 2 + i] = make_double2(b6[2 * i], b6[2 * i + 1]);
      res[rec] = commit.meta;  // the result word: final as it stands unless the loop pass takes the record
    }
  }
  // worklist of the loop pass: one atomic per warp (ticket[2] counts the entries; the loop's last CTA re-arms it)
  const uint32_t m = __ballot_sync(FULL, need);
  if (StageIn<IO>::on) {
    // prepared records out: whole (512 B) for the loop's QPs, lever arms and right-hand side (144 B) for the others --
    // none at all for records an EARLY set-up has finished
    const int lane = threadIdx.x & 31;
    const double* stage = stage_dyn + (threadIdx.x >> 5) * 32 * kStageStride;
    const int64_t base = (int64_t)blockIdx.x * kEdgeThreads + (threadIdx.x & ~31);
    const int cnt = n - base >= 32 ? 32 : (n > base ? (int)(n - base) : 0);
    __syncwarp();
    for (int e = 0; e < cnt; e++) {
      const int chunks = ((m >> e) & 1u) ? 32 : (EARLY ? 0 : 9);
      if (lane < chunks)
        reinterpret_cast<double2*>(prep + (base + e) * kPrepSize)[lane] = reinterpret_cast<const double2*>(stage + e * kStageStride)[lane];
    }
  }
  if (m) {
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(ticket + 2, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL, base, 0);
    if (need) work[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)rec;
  }
}

// ---- pass 2: the active-set loop ----------------------------------------------------------------------------------------
// Prepared records reach the warp through a staging buffer in shared memory, filled a batch (STAGE records) at a time by
// asynchronous copies (cp.async, 16 B per lane, one record per instruction) that are issued when the previous batch runs
// out and waited for only when a lane needs them; the worklist positions and record numbers of the batches after that
// are claimed (atomicAdd on the ticket) and loaded one and two batches ahead.  An idle lane therefore never sits on a
// chain of atomic -> index load -> record load (2 us, a whole loop iteration) as it did when every refill went to
// global memory on the spot (profiles/r02_ncu_tpq_v4_loop_cfg3_digest.txt: long-scoreboard stalls 1.96 per issue).


template <int LPQ>
__global__ void __launch_bounds__(LoopShape<LPQ>::THREADS, LoopShape<LPQ>::MIN_CTAS)
tpq_loop_kernel(const __grid_constant__ FastParams K, const double* __restrict__ prep, double* __restrict__ res,
                const uint32_t* __restrict__ work, unsigned long long* __restrict__ ticket) {
  constexpr int LPL = 4 / LPQ, NS = 32 / LPQ;  // legs per lane, QP slots per warp
  constexpr int STAGE = LoopShape<LPQ>::STAGE, kLoopThreads = LoopShape<LPQ>::THREADS;
  pdl_release();
  pdl_wait();  // the set-up pass is complete: worklist, prepared records and its counter are visible
  QPB_TL(0);
  __shared__ double side_all[(kLoopThreads / LPQ) * kSideSize];
  __shared__ __align__(16) double stage_all[(kLoopThreads / 32) * STAGE * kPrepSize];
  const int lane = threadIdx.x & 31;
  const int j = lane & (LPQ - 1);                             // lane within its QP
  const uint32_t leaders = 0xffffffffu / ((1u << LPQ) - 1u);  // bit of the first lane of every group
  const uint32_t glt = (1u << (lane & ~(LPQ - 1))) - 1u;      // lanes below this lane's group
  double* side = side_all + (threadIdx.x / LPQ) * kSideSize;
  double* stage = stage_all + (threadIdx.x >> 5) * STAGE * kPrepSize;
  constexpr int kRefill = (QPB_TPQ_REFILL + LPQ - 1) / LPQ;  // idle QP slots that trigger a retire + refill

  const int64_t n = (int64_t)*reinterpret_cast<volatile unsigned long long*>(ticket + 2);  // entries of the worklist
  // three batches in the pipe: A = staged (copies issued), B = record numbers loaded, C = worklist position claimed
  auto claim = [&]() -> int64_t {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(ticket, (unsigned long long)STAGE);
    return (int64_t)__shfl_sync(FULL, b, 0);
  };
  auto load_recs = [&](int64_t base) -> uint32_t { return (lane < STAGE && base + lane < n) ? __ldg(work + base + lane) : 0u; };
  int64_t baseB = claim();
  uint32_t recB = load_recs(baseB);
  int64_t baseC = claim();
  uint32_t recA = 0;    // lane i: record of staged entry i
  int st_n = 0, st_used = 0;  // staged entries of batch A and how many have been handed out
  bool st_flying = false;     // copies of batch A issued but not waited for yet
  auto issue_batch = [&]() {  // batch B -> A (start its copies), C -> B (load its record numbers), claim a new C
    const int64_t left = n - baseB;
    st_n = left <= 0 ? 0 : (left < STAGE ? (int)left : STAGE);
    st_used = 0;
    recA = recB;
    for (int e = 0; e < st_n; e++) {
      const uint32_t r = __shfl_sync(FULL, recA, e);
      cp_async16(stage + e * kPrepSize + 2 * lane, prep + (int64_t)r * kPrepSize + 2 * lane);
    }
    cp_async_commit();
    st_flying = st_n > 0;
    baseB = baseC;
    recB = load_recs(baseB);
    if (baseB < n) baseC = claim();
  };
  issue_batch();

  bool have = false;  // this lane's group holds a QP
  int64_t rec = 0;
  Lane<LPL> ln;
  {
    const double zero[12] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
    lane_init<LPL>(ln, j, zero, zero, zero, 0u, 0u, QPB_OK, 0, 0u);
    ln.done = true;
  }

  for (;;) {
    const uint32_t busy = __ballot_sync(FULL, have && !ln.done) & leaders;
    if (NS - __popc(busy) >= kRefill || busy == 0u) {
      // retire: a finished QP hands its final working set and iteration count to the finishing pass
      if (have && ln.done) {
        if (j == 0) res[rec] = pack_meta(ln.word, ln.stance, ln.status, ln.iters, 0u);
        have = false;
      }
      // refill: idle groups take the next staged records
      const uint32_t idle = ~__ballot_sync(FULL, have) & leaders;
      const int want = __popc(idle);
      if (want > 0 && st_used < st_n) {
        if (st_flying) {
          cp_async_wait_all();
          __syncwarp();
          st_flying = false;
        }
        const int slot = st_used + __popc(idle & glt);
        const bool take = !have && slot < st_n;
        const uint32_t r = __shfl_sync(FULL, recA, take ? slot : 0);
        if (take) {
          rec = (int64_t)r;
          const double* e = stage + slot * kPrepSize;
          const double meta = e[kPrepMeta];
          const uint32_t lo = (uint32_t)__double2loint(meta), hi = (uint32_t)__double2hiint(meta);
          // lane_init indexes all twelve with the lane's position in its QP
          lane_init<LPL>(ln, j, e + kPrepF, e + kPrepR, e + kPrepU, lo & 0xffffffu, (lo >> 24) & 15u, (int)((lo >> 28) & 3u),
                         (int)(hi & 0xffffu), ((hi >> 16) & 31u) | ((hi >> 21) << 31));
#pragma unroll
          for (int i = 0; i < (21 + LPQ - 1) / LPQ; i++) {
            const int k = j + LPQ * i;
            if (k < 21) side[kSideG + k] = e[kPrepG + k];
          }
#pragma unroll
          for (int i = 0; i < 12 / LPQ; i++) side[kSideR + j + LPQ * i] = e[kPrepR + j + LPQ * i];
          have = true;
        }
        // the slack of the first row: the lane that owns its leg has it, the others get it here
        const double slack = group_sum<LPQ>(take ? row_slack_share<LPL>(K, ln, j) : 0.0);
        if (take) ln.sp = slack;
        st_used += want < st_n - st_used ? want : st_n - st_used;
        __syncwarp();
      }
      if (st_used == st_n && baseB < n) issue_batch();  // start the next batch; lanes still idle pick it up a round later
      if (__ballot_sync(FULL, have) == 0u) {
        if (st_used == st_n) break;  // nothing held, nothing staged, nothing left to claim
        continue;
      }
    }
    iterate_group<LPQ>(K, ln, j, side);
    __syncwarp();  // G written by the first lane of a QP is read by its other lanes in the next round
  }
  // The last CTA out re-arms the work counter: the launch is self-contained, so the same counter slot serves graph
  // replays and later launches without a memset (ticket[0] = work counter, ticket[1] = CTAs finished).
  __syncthreads();
  QPB_TL(1);
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL) {
      ticket[3] = (unsigned long long)n;  // for a finishing pass that walks the worklist
      ticket[0] = 0ULL;
      ticket[1] = 0ULL;
      ticket[2] = 0ULL;  // the worklist counter the set-up pass filled
      __threadfence();
    }
  }
}

// ---- pass 3: polish + epilogue, one thread per record -------------------------------------------------------------------
// LIST: only the records of the worklist (ticket[3] entries, published by the loop pass) -- the others were finished by
// tpq_setup_kernel<IO, true>.
template <class IO, bool LIST>
__global__ void __launch_bounds__(kEdgeThreads, QPB_TPQ_FINISH_MINCTAS)
tpq_finish_kernel(const __grid_constant__ EdgeParams P, const __grid_constant__ FastParams K, IO io, int64_t n,
                  const double* __restrict__ prep, const double* __restrict__ res, const uint32_t* __restrict__ work,
                  const unsigned long long* __restrict__ ticket) {
  pdl_wait();  // the loop pass is complete: every result word is final
  int64_t rec = (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x;
  QPB_TL(2);
  if (LIST) {
    if (rec >= (int64_t)__ldg(ticket + 3)) return;
    rec = (int64_t)__ldg(work + rec);
  } else if (rec >= n) {
    return;
  }
  const double* e = prep + rec * kPrepSize;
  State st;
  double b6[6];
  const double w = __ldg(res + rec);
  const uint32_t lo = (uint32_t)__double2loint(w);
  st.word = lo & 0xffffffu;
  st.stance = (lo >> 24) & 15u;
  st.status = (int)((lo >> 28) & 3u);
  st.iters = __double2hiint(w) & 0xffff;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(e + kPrepR) + i);
    st.r[2 * i] = t.x;
    st.r[2 * i + 1] = t.y;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(e + kPrepB) + i);
    b6[2 * i] = t.x;
    b6[2 * i + 1] = t.y;
  }
#pragma unroll
  for (int i = 0; i < 12; i++) st.f[i] = st.u[i] = 0.0;
  double R[9], q[12], grf[12], tau[12];
  tpq_load_Rq(io, rec, R, q);
  bool qfin = true;  // the set-up checked slots 0..47; the joint angles are first touched here
#pragma unroll
  for (int i = 0; i < 12; i++) qfin = qfin && (fabs(q[i]) <= 1.79769313486231570e308);
  if (!qfin && st.status == QPB_OK) st.status = QPB_BAD_INPUT;
  polish(K, st, b6);  // the minimiser on the final faces, from scratch
  finish<4>(P, R, q, st, grf, tau);
  tpq_store(io, rec, grf, tau, st.status, st.iters, st.word | 0x80000000u);
  QPB_TL(3);
}

// ---- small batches: the three passes in ONE launch, one thread per record -------------------------------------------
// A per-tick caller (one robot, or a few hundred) waits for the answer, so what counts is the latency of a single QP,
// not throughput: one launch instead of three and no scratch memory; every record gets a warp to itself as long as
// there are warps to go round (records_per_cta = 1 up to 8 CTAs per SM), so no QP waits for another one's working-set
// changes.  The starting pair and R, q wait in shared memory instead of the prepared record; when the set-up's pair is
// already optimal (the usual outcome of a warm start) its forces ARE the minimiser on the final faces from one fresh
// solve, so the polish is skipped.  Same arithmetic as the three passes otherwise (tests compare the two).
constexpr int kOneThreads = 32;
enum : int { kKeepF = 0, kKeepU = 12, kKeepG = 24, kKeepR = 45, kKeepQ = 54, kKeepSize = 66 };

struct KeepCommit {
  double* e;  // this thread's column of the keep block (element i at e[i * kOneThreads])
  uint32_t key, word;
  int status, iters;
  __device__ __forceinline__ void operator()(const State& st, const double (&G)[21], uint32_t k) {
    key = k;
    word = st.word;
    status = st.status;
    iters = st.iters;
#pragma unroll
    for (int i = 0; i < 12; i++) e[(kKeepF + i) * kOneThreads] = st.f[i];
    if (k == 0u) return;  // optimal already: the loop will not run, its multipliers and G are not needed
#pragma unroll
    for (int i = 0; i < 12; i++) e[(kKeepU + i) * kOneThreads] = st.u[i];
#pragma unroll
    for (int i = 0; i < 21; i++) e[(kKeepG + i) * kOneThreads] = G[i];
  }
};

// flags (COOP only, may be null): completion words in pinned host memory, one per record, set to seq once the record's
// result is visible to the host -- the one-robot caller waits on them instead of synchronising the stream.
// COOP: one record per warp (records_per_cta = 1).  Lane 0 solves; the epilogue -- twelve sincos, the Jacobian columns,
// the store -- is spread over lanes 0..15 the way the half-warp kernel does it, a third of the instructions of the whole
// QP when one thread runs through them (a lone warp pays an instruction-cache miss for nearly every instruction it
// executes once: profiles/r02_ncu_one_n1_digest.txt).
template <class IO, bool COOP>
__global__ void __launch_bounds__(kOneThreads)
tpq_one_kernel(const __grid_constant__ EdgeParams P, const __grid_constant__ FastParams K, IO io, int64_t n, int records_per_cta,
               uint32_t* __restrict__ flags, uint32_t seq) {
  __shared__ double side_all[kOneThreads * kSideSize];
  __shared__ double keep_all[kOneThreads * kKeepSize];
  const int t = threadIdx.x;
  const int64_t rec = COOP ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * records_per_cta + t;
  const bool active = COOP ? t == 0 : (t < records_per_cta && rec < n);
  double* side = side_all + t * kSideSize;
  double* keep = keep_all + t;
  State st;
  double b6[6];
  Lane<4> ln;
  ln.done = true;
  st.status = QPB_OK;
  st.iters = 0;
  st.word = st.stance = 0u;
  bool looped = false, qfin = true;
  uint32_t key = 0u;  // first row of the loop (0: the set-up's pair is optimal already)
  if (active) {
    double v[48];
    uint32_t cbytes, hint;
    tpq_load(io, rec, v, cbytes, hint);
#pragma unroll
    for (int i = 0; i < 9; i++) keep[(kKeepR + i) * kOneThreads] = v[kR + i];
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const double qi = tpq_q(io, rec, v, i);
      qfin = qfin && (fabs(qi) <= 1.79769313486231570e308);  // the set-up checks slots 0..47, the joint angles are checked here
      keep[(kKeepQ + i) * kOneThreads] = qi;
    }
    double G[21];
    KeepCommit commit{ keep, 0u, 0u, QPB_OK, 0 };
    setup(P, K, v, cbytes, hint, st, b6, G, commit);
    // the last pair committed is where the solve stands (a later block round may have been tried and rejected)
    st.word = commit.word;
    st.status = commit.status;
    st.iters = commit.iters;
    key = commit.key;
    looped = key != 0u;
#pragma unroll
    for (int i = 0; i < 12; i++) st.f[i] = keep[(kKeepF + i) * kOneThreads];
    if (looped) {
#pragma unroll
      for (int i = 0; i < 12; i++) st.u[i] = keep[(kKeepU + i) * kOneThreads];
    }
  }
  if (__any_sync(FULL, looped)) {
    // every lane of the warp walks through the loop (finished and idle ones with their stores predicated off)
    const double zero[12] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
    lane_init<4>(ln, 0, zero, zero, zero, 0u, 0u, QPB_OK, 0, 0u);
    ln.done = true;
    if (looped) {
      lane_init<4>(ln, 0, st.f, st.r, st.u, st.word, st.stance, st.status, st.iters, key);
#pragma unroll
      for (int i = 0; i < 21; i++) side[kSideG + i] = keep[(kKeepG + i) * kOneThreads];
#pragma unroll
      for (int i = 0; i < 12; i++) side[kSideR + i] = st.r[i];
      ln.sp = row_slack_share<4>(K, ln, 0);
    }
    while (__any_sync(FULL, !ln.done)) iterate_group<1>(K, ln, 0, side);
    if (looped) {
      st.word = ln.word;
      st.status = ln.status;
      st.iters = ln.iters;
    }
  }
  if (!qfin && st.status == QPB_OK) st.status = QPB_BAD_INPUT;
  if (looped) polish(K, st, b6);  // (a pair the set-up found optimal IS the minimiser on its faces from one fresh solve)
  if (COOP) {
    if (active) {
#pragma unroll
      for (int i = 0; i < 12; i++) keep[(kKeepF + i) * kOneThreads] = st.f[i];
    }
    __syncwarp();
    const int status = __shfl_sync(FULL, st.status, 0);
    const int iters = __shfl_sync(FULL, st.iters, 0);
    const uint32_t word = __shfl_sync(FULL, st.word, 0), stance = __shfl_sync(FULL, st.stance, 0);
    const int vi = t < 12 ? t : 0, leg = vi / 3, ax = vi - 3 * leg;
    const double* kp = keep_all;  // the record of lane 0: element i at kp[i * kOneThreads]
    const double qraw = kp[(kKeepQ + vi) * kOneThreads];
    const bool on = status == QPB_OK && ((stance >> leg) & 1u), qok = status != QPB_BAD_INPUT;
    const double f0 = kp[(kKeepF + 3 * leg) * kOneThreads], f1 = kp[(kKeepF + 3 * leg + 1) * kOneThreads],
                 f2 = kp[(kKeepF + 3 * leg + 2) * kOneThreads];
    const double fb = on ? -1.0 * (kp[(kKeepR + ax) * kOneThreads] * f0 + kp[(kKeepR + 3 + ax) * kOneThreads] * f1 +
                                   kp[(kKeepR + 6 + ax) * kOneThreads] * f2)
                         : 0.0;
    const double qa = qraw + (ax == 2 ? kp[(kKeepQ + 3 * leg + 1) * kOneThreads] : 0.0);  // t1, t2, t2 + t3
    double sn, cs;
    sincos(qok ? qa : 0.0, &sn, &cs);
    const double s1 = __shfl_sync(FULL, sn, 3 * leg), c1 = __shfl_sync(FULL, cs, 3 * leg);
    const double s2 = __shfl_sync(FULL, sn, 3 * leg + 1), c2 = __shfl_sync(FULL, cs, 3 * leg + 1);
    const double s23 = __shfl_sync(FULL, sn, 3 * leg + 2), c23 = __shfl_sync(FULL, cs, 3 * leg + 2);
    const double fbx = __shfl_sync(FULL, fb, 3 * leg), fby = __shfl_sync(FULL, fb, 3 * leg + 1), fbz = __shfl_sync(FULL, fb, 3 * leg + 2);
    double Jx, Jy, Jz;
    leg_jacobian_col(ax, P.link[3 * leg], P.link[3 * leg + 1], P.link[3 * leg + 2], s1, c1, s2, c2, s23, c23, Jx, Jy, Jz);
    double tau = Jx * fbx + Jy * fby + Jz * fbz;
    if (P.clamp_tau) tau = fmin(fmax(tau, P.tau_min), P.tau_max);
    if (!on) tau = 0.0;
    store_rec(io, rec, t, fb, tau, status, iters, word | 0x80000000u);
    if (flags) {
      // a host thread is spinning on flags[rec] (pinned memory): the record first, system-wide, then the flag
      __syncwarp();
      if (t == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(flags + rec) = seq;
      }
    }
    return;
  }
  if (!active) return;
  // The epilogue of finish() rolled over the legs (one copy of the sincos code: the loop above shares the instruction
  // cache with it), its per-leg arrays in this thread's column of shared memory -- as local-memory arrays they miss the
  // L1 that 8 CTAs x 26 KB of shared memory leave (profiles/r02_finish_unroll_ab.txt).
  double R[9], grf[12], tau[12];
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = keep[(kKeepR + i) * kOneThreads];
#pragma unroll
  for (int i = 0; i < 12; i++) keep[(kKeepF + i) * kOneThreads] = st.f[i];
  finish_rolled_smem<kKeepF, kKeepQ, kKeepU, kKeepG>(P, R, keep, kOneThreads, st.status, st.stance);  // (U and G are free by now)
#pragma unroll
  for (int i = 0; i < 12; i++) {
    grf[i] = keep[(kKeepU + i) * kOneThreads];
    tau[i] = keep[(kKeepG + i) * kOneThreads];
  }
  tpq_store(io, rec, grf, tau, st.status, st.iters, st.word | 0x80000000u);
}

}  // namespace tpq
}  // namespace qpb
