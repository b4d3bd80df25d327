// qpb_api.cu -- C ABI of libqpb200.so (include/qpb200.h) on top of the kernels in qpb_kernel.cuh.
// Host side of the drop-in boundary: plain pointers in, plain pointers out, no exceptions.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "qpb_internal.h"
#include "qpb_kernel.cuh"
#include "qpb_kernel16.cuh"
#include "qpb_plan.cuh"
#include "qpb_swing.cuh"
#include "qpb_tpq.cuh"
#include "qpb_wire.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define QPB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      return fail(QPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
    }                                                                                          \
  } while (0)

// dense Cholesky test for "symmetric positive definite"
bool is_sympd(const double* A, int n) {
  double L[144];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      const double a = A[i * n + j], b = A[j * n + i];
      if (std::fabs(a - b) > 1e-12 * (std::fabs(a) + std::fabs(b))) return false;
    }
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (!(s > 0.0)) return false;
    L[j * n + j] = std::sqrt(s);
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / L[j * n + j];
    }
  }
  return true;
}

constexpr int kHostSlots = 6;          // streams / staging slots of the host-buffer pipeline
constexpr int64_t kHostChunkMax = 16384;  // capacity of a pipeline stage (8 MiB in, 4 MiB out)
constexpr int64_t kMaxRecordsPerLaunch = (int64_t)1 << 31;  // the kernels index work items with 32 bits
constexpr uint32_t kTicketSlots = 4096;  // ring of {work counter, CTAs finished} pairs; the kernels re-arm their own pair
constexpr int64_t kSmallCall = 128;      // host calls up to this many robots go through one pinned, mapped staging block
constexpr int64_t kTpqChunk = (int64_t)1 << 20;  // records per set-up / loop / finish triple (512 MiB of scratch)
constexpr size_t kSmallFlagsOff = (size_t)kSmallCall * 1536;  // states + results + swing records (or the kinematics arrays)
constexpr size_t kSmallBytes = kSmallFlagsOff + (size_t)kSmallCall * sizeof(uint32_t);  // + one completion word per record

}  // namespace

int qpb_internal_fail(int code, const std::string& msg) { return fail(code, msg); }

struct qpb_handle {
  int device = 0;
  int num_sms = 0;
  int ctas_per_sm_packed = 0, ctas_per_sm_split = 0, ctas_per_sm_16 = 0;
  // kernel mapping: 32 = balance_qp_tpq_kernel (one thread per QP, range-space form; the default whenever W = w I and
  // fzmin >= 0, i.e. for the reference's configuration), 2 = balance_qp_kernel16 (two QPs per warp; general W),
  // 1 = balance_qp_kernel (one warp per QP, north_star's literal mapping).  QPB_QPS_PER_WARP=1|2|32 in the environment
  // at qpb_create time overrides the choice (32 is refused when the parameters do not qualify).
  int qps_per_warp = 2;
  int ctas_per_sm_tpq[5] = {};  // indexed by lanes per QP (1, 2, 4)
  int tpq_lpq = 1;              // lanes per QP of the range-space loop kernel (QPB_TPQ_LPQ=1|2|4 overrides)
  int tpq_pdl = 1;              // loop and finishing kernels of the three-pass path as programmatic dependent launches
                                // (qpb_tpq.cuh, pdl_wait; QPB_TPQ_PDL=0: plain launches)
  int64_t tpq_min_n = 12288;    // smaller batches take a one-launch kernel: lower latency (QPB_TPQ_MIN_N)
  int64_t tpq_one_max = 1;      // ... up to here the range-space one (tpq_one_kernel), above it the half-warp kernel (QPB_TPQ_ONE_MAX)
  int warm_batches = 0;         // device-resident calls: the records carry warm-start words (qpb_set_warm_batches)
  int64_t tpq_warm_defer_min = 196608;  // warm batches below this size take tpq_one_kernel, larger ones the three passes
                                        // with early finish (QPB_TPQ_WARM_DEFER_MIN; profiles/r02_warm_sizes.txt)
  qpb::tpq::FastParams fast;
  qpb::tpq::EdgeParams edge;  // the slice of params the set-up / finishing passes take by value
  qpb_params params;
  qpb_params* d_params = nullptr;
  cudaStream_t streams[kHostSlots] = {};  // per-slot compute streams of the host pipeline ([0] also serves the small calls)
  // All uploads of the pipeline go down ONE stream and all downloads down another, so the copy engines get their work
  // back to back (copies issued from six different streams left a 5-us gap between consecutive uploads); events carry
  // the dependencies: upload -> kernels -> download, and a slot's buffers are refilled only after their last reader.
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  cudaEvent_t ev_up[kHostSlots] = {}, ev_in_free[kHostSlots] = {}, ev_solved[kHostSlots] = {}, ev_down[kHostSlots] = {};
  qpb_state_rec* d_in[kHostSlots] = {};
  qpb_out_rec* d_out[kHostSlots] = {};
  qpb_swing_rec* d_sw[kHostSlots] = {};
  qpb_wire_state* d_win[kHostSlots] = {};  // wire records of a stage as they arrive / leave (qpb_control_batch_wire_host)
  qpb_wire_out* d_wout[kHostSlots] = {};
  double* d_scratch[kHostSlots] = {};  // scratch of the three-pass path for one pipeline stage (instead of cudaMallocAsync)
  // QPB_HOST_TRACE=1: every stage of the host pipeline is bracketed by events (stream reached the stage, upload done,
  // kernels done, download done); the next synchronisation prints the timeline to stderr.  A debugging aid.
  int trace = 0;
  struct TraceStage { cudaEvent_t ev[4]; int slot; int64_t m; };
  std::vector<TraceStage> trace_stages;
  int host_cstreams = 2;               // compute streams the stages rotate over: 2 measured best, 3 worst (QPB_HOST_CSTREAMS,
                                       // profiles/r02_host_pipeline_streams.txt)
  int host_stages = 8;                 // stages a host batch is cut into on the three-pass path (QPB_HOST_STAGES)
  qpb_joint_gains* d_gains = nullptr;  // JointController gains for the swing-leg half of the tick
  qpb_plan_params* d_plan = nullptr;   // FootPlanner / FootTrajectoryManager constants
  std::atomic<int64_t> launches{ 0 };
  // ring of work-ticket counters, one per in-flight launch of the balance kernel
  unsigned long long* d_tickets = nullptr;
  std::atomic<uint32_t> ticket_slot{ 0 };
  // Per-tick callers (the C++ shim: one robot, pageable stack records) get a latency path: the records are copied into
  // a pinned, mapped block the kernel reads and writes over PCIe itself -- one launch and one stream synchronisation,
  // no cudaMemcpy, no cudaMalloc.
  unsigned char* h_small = nullptr;  // host address of the pinned block
  unsigned char* d_small = nullptr;  // its device alias
  int64_t host_chunk = 8192;  // records per H2D/kernel/D2H pipeline stage (QPB_HOST_CHUNK overrides)
  uint64_t host_slot = 0;     // next stage of the host pipeline (stages of successive asynchronous calls keep rotating)
  uint32_t small_seq = 0;     // completion stamp of the latency path's last call
  int small_poll = 1;         // the latency path waits on completion words in pinned memory (QPB_SMALL_POLL=0: stream sync)
  int zero_copy = 1;          // pinned host buffers are read/written by the kernel itself over PCIe (QPB_ZEROCOPY=0: always stage)
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

inline qpb::PackedIO offset_io(const qpb::PackedIO& io, int64_t lo) { return qpb::PackedIO{ io.in + lo, io.out + lo }; }
inline qpb::SplitIO offset_io(const qpb::SplitIO& io, int64_t lo) {
  return qpb::SplitIO{ io.Rwb + 9 * lo, io.Rwb_d + 9 * lo, io.x + 3 * lo, io.xdot + 3 * lo, io.w + 3 * lo, io.x_d + 3 * lo,
                       io.xdot_d + 3 * lo, io.w_d + 3 * lo, io.feet + 12 * lo, io.q + 12 * lo, io.contact + 4 * lo,
                       io.grf + 12 * lo, io.tau ? io.tau + 12 * lo : nullptr, io.status ? io.status + lo : nullptr };
}

template <class IO>
int launch_balance(qpb_handle* h, const IO& io, int64_t n, int ctas_per_sm, cudaStream_t stream, int force_path = 0,
                   uint32_t* flags = nullptr, uint32_t seq = 0, int64_t whole_n = 0, double* scratch = nullptr) {
  if (n == 0) return QPB_SUCCESS;
  if (n > kMaxRecordsPerLaunch) return fail(QPB_ERR_INVALID_ARG, "more than 2^31 records in one call: split the batch");
  // 1: one warp per QP; 2: two QPs per warp (half-warp kernel); 32: range-space path.  force_path: a caller that cuts a
  // batch into pieces (the host pipeline) picks the path once, by the size of the whole batch, so that a record's result
  // does not depend on which piece it fell into.
  int per_warp = h->qps_per_warp;
  // Launch i draws its work tickets from pair i % R of the ring; the last CTA of a launch re-arms the pair, so a launch
  // is self-contained (safe under CUDA-graph replay) as long as fewer than R = 4096 launches of a handle are in flight.
  const uint32_t slot = h->ticket_slot.fetch_add(1, std::memory_order_relaxed) % kTicketSlots;
  unsigned long long* t0 = h->d_tickets + 4 * (size_t)slot;  // {work ticket, CTAs finished, worklist entries, -}
  // A warm batch (records carrying last tick's working sets) is at its optimum after the set-up's first solve almost
  // always, so the one-launch kernel wins at every size: no scratch, no second and third pass.
  // Warm batches: one launch (33) up to ~200 000 records.  Larger ones: three passes in which the set-up finishes the
  // records that are optimal at once and the loop and finishing passes only see the rest (34) -- in one launch nearly
  // every warp holds at least one record that needs the loop and waits for it, which costs more than two extra
  // launches once the batch is large (1 048 576 records one tick later: 438 us against 520 us; 65 536: 65 against 52).
  if (per_warp == 32 && !force_path && h->warm_batches && std::is_same<IO, qpb::PackedIO>::value)
    force_path = n < h->tpq_warm_defer_min ? 33 : 34;
  if (per_warp == 32 && (force_path ? (force_path == 32 || force_path == 34) : n >= h->tpq_min_n)) {
    // Three passes over scratch memory (qpb_tpq.cuh): set-up -> prepared records, the active-set loop, polish + epilogue.
    // The scratch comes from the stream-ordered allocator, so concurrent calls on different streams never share it.
    const int lpq = h->tpq_lpq;
    bool first_chain = true;
    // one chain of three launches per at most 2^20 records of [lo0, lo0 + n0), on stream s
    // warm batch: records optimal after the set-up are finished there (for cold batches -- 18 % / 45 % such records on
    // config 2 / 3 -- the same costs 7 %: every warp pays the epilogue twice; profiles/r02_early_cold_ab.txt)
    const bool early = force_path == 34;
    auto run_chain = [&](int64_t lo0, int64_t n0, cudaStream_t s) -> int {
      for (int64_t lo = lo0; lo < lo0 + n0; lo += kTpqChunk) {
        const int64_t m = lo0 + n0 - lo < kTpqChunk ? lo0 + n0 - lo : kTpqChunk;
        const IO part = offset_io(io, lo);
        double* prep = scratch;  // m prepared records, m result words, the worklist of the loop pass (m record indices)
        const size_t prep_bytes = (size_t)m * qpb::tpq::kPrepSize * sizeof(double);
        if (!scratch)  // (the host pipeline brings its own: one block per stage slot, n <= kHostChunkMax)
          QPB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&prep), prep_bytes + (size_t)m * (sizeof(double) + sizeof(uint32_t)), s));
        double* res = prep + (size_t)m * qpb::tpq::kPrepSize;
        uint32_t* work = reinterpret_cast<uint32_t*>(res + m);
        const uint32_t slot2 = first_chain ? slot : h->ticket_slot.fetch_add(1, std::memory_order_relaxed) % kTicketSlots;
        first_chain = false;
        unsigned long long* tk = h->d_tickets + 4 * (size_t)slot2;
        const unsigned edge = (unsigned)((m + qpb::tpq::kEdgeThreads - 1) / qpb::tpq::kEdgeThreads);
        const size_t stage_bytes = qpb::tpq::StageIn<IO>::on ? qpb::tpq::kSetupStageBytes : 0;
        if (early)
          qpb::tpq::tpq_setup_kernel<IO, true><<<edge, qpb::tpq::kEdgeThreads, stage_bytes, s>>>(h->edge, h->fast, part, m, prep, res, work, tk);
        else
          qpb::tpq::tpq_setup_kernel<IO, false><<<edge, qpb::tpq::kEdgeThreads, stage_bytes, s>>>(h->edge, h->fast, part, m, prep, res, work, tk);
        // Programmatic dependent launches for the second and third pass of a device-resident call.  Not while the stream is
        // being captured (a graph keeps the plain edges), and not for the stages of the host pipeline (they bring their own
        // scratch): there the stages of two compute streams fill each other's idle SMs, and CTAs parked in griddepcontrol.wait
        // hold the registers the other stream's kernels would have used -- end to end 1.005e8 -> 9.2e7 QP/s with it
        // (profiles/r02_pdl_ab.txt).
        cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
        const bool pdl = h->tpq_pdl && !scratch && cudaStreamIsCapturing(s, &cap_status) == cudaSuccess && cap_status == cudaStreamCaptureStatusNone;
        cudaLaunchAttribute pdl_attr[1];
        pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
        auto launch_cfg = [&](unsigned grid_x, int threads) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(grid_x);
          cfg.blockDim = dim3((unsigned)threads);
          cfg.stream = s;
          cfg.attrs = pdl_attr;
          cfg.numAttrs = pdl ? 1 : 0;
          return cfg;
        };
        const int lthreads = lpq == 1 ? qpb::tpq::LoopShape<1>::THREADS : (lpq == 2 ? qpb::tpq::LoopShape<2>::THREADS : qpb::tpq::LoopShape<4>::THREADS);
        const int64_t want = (m * lpq + lthreads - 1) / lthreads;
        const int64_t cap = (int64_t)h->num_sms * h->ctas_per_sm_tpq[lpq];
        const int grid = (int)(want < cap ? want : cap);
        {
          const cudaLaunchConfig_t cfg = launch_cfg((unsigned)grid, lthreads);
          const double* cprep = prep;
          const uint32_t* cwork = work;
          if (lpq == 1)
            QPB_CUDA(cudaLaunchKernelEx(&cfg, qpb::tpq::tpq_loop_kernel<1>, h->fast, cprep, res, cwork, tk));
          else if (lpq == 2)
            QPB_CUDA(cudaLaunchKernelEx(&cfg, qpb::tpq::tpq_loop_kernel<2>, h->fast, cprep, res, cwork, tk));
          else
            QPB_CUDA(cudaLaunchKernelEx(&cfg, qpb::tpq::tpq_loop_kernel<4>, h->fast, cprep, res, cwork, tk));
        }
        {
          const cudaLaunchConfig_t cfg = launch_cfg(edge, qpb::tpq::kEdgeThreads);
          const double *cprep = prep, *cres = res;
          const uint32_t* cwork = work;
          const unsigned long long* ctk = tk;
          if (early)
            QPB_CUDA(cudaLaunchKernelEx(&cfg, qpb::tpq::tpq_finish_kernel<IO, true>, h->edge, h->fast, part, m, cprep, cres, cwork, ctk));
          else
            QPB_CUDA(cudaLaunchKernelEx(&cfg, qpb::tpq::tpq_finish_kernel<IO, false>, h->edge, h->fast, part, m, cprep, cres, cwork, ctk));
        }
        h->launches.fetch_add(3, std::memory_order_relaxed);
        QPB_CUDA(cudaGetLastError());
        if (!scratch) QPB_CUDA(cudaFreeAsync(prep, s));
      }
      return QPB_SUCCESS;
    };
    // (Cutting a large batch into parts whose chains run side by side on helper streams was measured and dropped: +1 % with
    // two parts on config 2, a loss everywhere else -- profiles/r02_split_ab.txt.)
    return run_chain(0, n, stream);
  }
  if (per_warp == 32 && (force_path ? force_path == 33 : n <= h->tpq_one_max)) {
    // small batches: set-up, loop and finish in one launch, a warp per record while there are warps to go round
    int64_t per_cta = (n + (int64_t)h->num_sms * 8 - 1) / ((int64_t)h->num_sms * 8);
    if (per_cta > qpb::tpq::kOneThreads) per_cta = qpb::tpq::kOneThreads;
    const unsigned grid = (unsigned)((n + per_cta - 1) / per_cta);
    // a warp per record: the epilogue is shared out over its lanes (decided on the batch a shard is cut from: the two
    // epilogues round differently in the last bit)
    if (per_cta == 1 && (whole_n > n ? whole_n : n) <= (int64_t)h->num_sms * 8)
      qpb::tpq::tpq_one_kernel<IO, true><<<grid, qpb::tpq::kOneThreads, 0, stream>>>(h->edge, h->fast, io, n, 1, flags, seq);
    else
      qpb::tpq::tpq_one_kernel<IO, false><<<grid, qpb::tpq::kOneThreads, 0, stream>>>(h->edge, h->fast, io, n, (int)per_cta, nullptr, 0u);
    h->launches.fetch_add(1, std::memory_order_relaxed);
    QPB_CUDA(cudaGetLastError());
    return QPB_SUCCESS;
  }
  if (per_warp == 32) per_warp = 2;  // in between: the half-warp kernel (one launch, 16 lanes per QP)
  const int64_t units = (n + per_warp - 1) / per_warp;
  const int64_t want = (units + qpb::WARPS_PER_CTA - 1) / qpb::WARPS_PER_CTA;
  const int64_t cap = (int64_t)h->num_sms * (per_warp == 2 ? h->ctas_per_sm_16 : ctas_per_sm);
  const int grid = (int)(want < cap ? want : cap);
  if (per_warp == 2) {
    const qpb_params& p = h->params;
    const qpb::LoopConsts kc{ p.mu, p.fzmin, p.fzmax, -1e-9 * (1.0 + std::fmax(std::fabs(p.fzmin), std::fabs(p.fzmax))), p.max_iter };
    qpb::balance_qp_kernel16<IO><<<grid, qpb::WARPS_PER_CTA * 32, 0, stream>>>(h->d_params, io, n, t0, kc);
  }
  else
    qpb::balance_qp_kernel<IO><<<grid, qpb::WARPS_PER_CTA * 32, 0, stream>>>(h->d_params, io, n, t0);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  QPB_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

// swing-leg kernel after the balance kernel on the same stream (stream order = the merge of commander_node.cpp:515)
int launch_swing(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, const qpb_swing_rec* d_swing, qpb_out_rec* d_out,
                 cudaStream_t stream) {
  if (n == 0) return QPB_SUCCESS;
  const int threads = 128;
  qpb::swing_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(h->d_params, h->d_gains, d_states,
                                                                                    d_swing, d_out, n);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  QPB_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

int ensure_small(qpb_handle* h) {
  if (h->h_small) return QPB_SUCCESS;
  void* p = nullptr;
  QPB_CUDA(cudaHostAlloc(&p, kSmallBytes, cudaHostAllocMapped));
  std::memset(p, 0, kSmallBytes);
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) {
    cudaFreeHost(p);
    return fail(QPB_ERR_CUDA, "cudaHostGetDevicePointer failed");
  }
  h->h_small = static_cast<unsigned char*>(p);
  h->d_small = static_cast<unsigned char*>(d);
  if (!h->streams[0]) QPB_CUDA(cudaStreamCreateWithFlags(&h->streams[0], cudaStreamNonBlocking));
  return QPB_SUCCESS;
}

// Host-buffer path shared by qpb_control_batch_host and qpb_tick_batch_host.  Pinned buffers of the balance path are
// handed to the kernel directly; otherwise stages of records are uploaded, solved and downloaded on a ring of
// streams so the three overlap (the tick always stages: its second kernel would re-read the records over PCIe).
int sync_pipeline(qpb_handle* h) {
  cudaError_t first = cudaSuccess;
  for (int s = 0; s < kHostSlots + 2; s++) {
    cudaStream_t st = s < kHostSlots ? h->streams[s] : (s == kHostSlots ? h->up_stream : h->down_stream);
    if (st) {
      const cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess && first == cudaSuccess) first = e;
    }
  }
  if (first != cudaSuccess) return fail(QPB_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(first));
  if (h->trace && !h->trace_stages.empty()) {
    std::fprintf(stderr, "qpb host pipeline: %zu stages (us since the first one was reached)\n  #  slot records   reached  uploaded    solved downloaded\n",
                 h->trace_stages.size());
    for (size_t i = 0; i < h->trace_stages.size(); i++) {
      float t[4] = { 0.f, 0.f, 0.f, 0.f };
      for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], h->trace_stages[0].ev[0], h->trace_stages[i].ev[k]);
      std::fprintf(stderr, "%3zu  %4d %7lld %9.1f %9.1f %9.1f %9.1f\n", i, h->trace_stages[i].slot, (long long)h->trace_stages[i].m,
                   t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3);
    }
    for (auto& ts : h->trace_stages)
      for (int k = 0; k < 4; k++) cudaEventDestroy(ts.ev[k]);
    (void)cudaGetLastError();
    h->trace_stages.clear();
  }
  return QPB_SUCCESS;
}

// whole_n / whole_warm: when the n records are one shard of a larger batch (qpb_multi_*), the size of that batch and
// whether ITS first record carries a warm-start word -- the kernels are chosen for the batch, not for the shard.
// w_in / w_out (instead of h_states / h_out): the batch travels as wire records and is widened / narrowed on the device.
int host_pipeline(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, const qpb_swing_rec* h_swing, qpb_out_rec* h_out,
                  bool async = false, int64_t whole_n = -1, int whole_warm = -1, const qpb_wire_state* w_in = nullptr,
                  qpb_wire_out* w_out = nullptr) {
  if (n == 0) return QPB_SUCCESS;
  const bool wire = w_in != nullptr;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  // one path for the whole batch, whatever the pieces it is cut into
  // (host records can be looked at: a batch whose first record carries a warm-start word is taken to be a warm batch)
  const int64_t dn = whole_n >= 0 ? whole_n : n;
  const bool hinted = whole_warm >= 0 ? whole_warm != 0 : (wire ? (w_in[0].warm >> 31) != 0u : (h_states[0].pad[3] & 0x80u) != 0);
  const bool warm = h->qps_per_warp == 32 && (h->warm_batches || hinted);
  const bool range_space = h->qps_per_warp == 32 && (warm ? dn >= h->tpq_warm_defer_min : dn >= h->tpq_min_n);
  const int path = range_space ? (warm ? 34 : 32) : (h->qps_per_warp == 32 ? ((warm || dn <= h->tpq_one_max) ? 33 : 2) : h->qps_per_warp);
  if (!async && h->zero_copy && n <= kSmallCall) {
    // Latency path for per-tick callers: stage through the handle's pinned block, kernels work on its device alias.
    const int rc0 = ensure_small(h);
    if (rc0 != QPB_SUCCESS) return rc0;
    const size_t off_out = (size_t)kSmallCall * sizeof(qpb_state_rec), off_sw = off_out + (size_t)kSmallCall * sizeof(qpb_out_rec);
    if (wire) {  // a handful of records: widen them on the way into the pinned block
      std::memset(h->h_small, 0, (size_t)n * sizeof(qpb_state_rec));
      for (int64_t i = 0; i < n; i++) std::memcpy(h->h_small + (size_t)i * sizeof(qpb_state_rec), w_in + i, sizeof(qpb_wire_state));
    } else {
      std::memcpy(h->h_small, h_states, (size_t)n * sizeof(qpb_state_rec));
    }
    if (h_swing) std::memcpy(h->h_small + off_sw, h_swing, (size_t)n * sizeof(qpb_swing_rec));
    const qpb_state_rec* ds = reinterpret_cast<const qpb_state_rec*>(h->d_small);
    qpb_out_rec* dout = reinterpret_cast<qpb_out_rec*>(h->d_small + off_out);
    qpb::PackedIO io{ ds, dout };
    // One-launch range-space kernel, a warp per record: each warp stamps a completion word in the pinned block once its
    // result is visible to the host, and the host waits on those words instead of synchronising the stream (1.3 us less
    // per call on the B200 box, profiles/r02_launch_floor.txt).  Anything else, or no stamp within 2 ms: stream sync.
    const bool poll = h->small_poll && path == 33 && !h_swing && dn <= (int64_t)h->num_sms * 8;
    uint32_t seq = 0;
    if (poll) {
      if (++h->small_seq == 0u) h->small_seq = 1u;
      seq = h->small_seq;
    }
    int rc = launch_balance(h, io, n, h->ctas_per_sm_packed, h->streams[0], path,
                            poll ? reinterpret_cast<uint32_t*>(h->d_small + kSmallFlagsOff) : nullptr, seq, dn);
    if (rc == QPB_SUCCESS && h_swing)
      rc = launch_swing(h, n, ds, reinterpret_cast<const qpb_swing_rec*>(h->d_small + off_sw), dout, h->streams[0]);
    if (rc != QPB_SUCCESS) return rc;
    bool stamped = false;
    if (poll) {
      const volatile uint32_t* flags = reinterpret_cast<const volatile uint32_t*>(h->h_small + kSmallFlagsOff);
      const auto t0 = std::chrono::steady_clock::now();
      stamped = true;
      uint32_t spins = 0;
      for (int64_t i = 0; i < n && stamped; i++)
        while (flags[i] != seq)
          if ((++spins & 0x3ffu) == 0u && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) {
            stamped = false;
            break;
          }
      std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (!stamped) QPB_CUDA(cudaStreamSynchronize(h->streams[0]));
    if (wire) {
      for (int64_t i = 0; i < n; i++) {
        const qpb_out_rec* o = reinterpret_cast<const qpb_out_rec*>(h->h_small + off_out) + i;
        std::memcpy(w_out + i, o, 24 * sizeof(double));
        w_out[i].status = (int16_t)o->status;
        w_out[i].iters = (int16_t)(o->iters < 0 ? 0 : (o->iters > 32767 ? 32767 : o->iters));
        std::memcpy(&w_out[i].wset, o->pad, sizeof(uint32_t));
      }
    } else {
      std::memcpy(h_out, h->h_small + off_out, (size_t)n * sizeof(qpb_out_rec));
    }
    return QPB_SUCCESS;
  }
  if (!async && h->zero_copy && !h_swing && !range_space && !wire) {
    // Pinned (hence mapped) buffers and the one-launch kernels: the launch reads the records and writes the results
    // straight over PCIe, one coalesced 512-B request per record.  No staging copies, no pipeline fill/drain.  (The
    // range-space path makes three passes over its records, so it always stages them in device memory.)
    cudaPointerAttributes ai, ao;
    if (cudaPointerGetAttributes(&ai, h_states) == cudaSuccess && cudaPointerGetAttributes(&ao, h_out) == cudaSuccess &&
        ai.type == cudaMemoryTypeHost && ao.type == cudaMemoryTypeHost && ai.devicePointer && ao.devicePointer) {
      if (!h->streams[0]) QPB_CUDA(cudaStreamCreateWithFlags(&h->streams[0], cudaStreamNonBlocking));
      qpb::PackedIO io{ static_cast<const qpb_state_rec*>(ai.devicePointer), static_cast<qpb_out_rec*>(ao.devicePointer) };
      const int rc = launch_balance(h, io, n, h->ctas_per_sm_packed, h->streams[0], path, nullptr, 0u, dn);
      if (rc != QPB_SUCCESS) return rc;
      QPB_CUDA(cudaStreamSynchronize(h->streams[0]));
      return QPB_SUCCESS;
    } else {
      (void)cudaGetLastError();
    }
  }
  if (!h->up_stream) QPB_CUDA(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
  if (!h->down_stream) QPB_CUDA(cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
  for (int s = 0; s < kHostSlots; s++) {  // lazily create the pipeline
    if (!h->streams[s]) QPB_CUDA(cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking));
    for (cudaEvent_t* e : { &h->ev_up[s], &h->ev_in_free[s], &h->ev_solved[s], &h->ev_down[s] })
      if (!*e) QPB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (!h->d_in[s]) QPB_CUDA(cudaMalloc(&h->d_in[s], kHostChunkMax * sizeof(qpb_state_rec)));
    if (!h->d_out[s]) QPB_CUDA(cudaMalloc(&h->d_out[s], kHostChunkMax * sizeof(qpb_out_rec)));
    if (h_swing && !h->d_sw[s]) QPB_CUDA(cudaMalloc(&h->d_sw[s], kHostChunkMax * sizeof(qpb_swing_rec)));
    if (wire && !h->d_win[s]) QPB_CUDA(cudaMalloc(&h->d_win[s], kHostChunkMax * sizeof(qpb_wire_state)));
    if (wire && !h->d_wout[s]) QPB_CUDA(cudaMalloc(&h->d_wout[s], kHostChunkMax * sizeof(qpb_wire_out)));
    if (range_space && !h->d_scratch[s])
      QPB_CUDA(cudaMalloc(&h->d_scratch[s], kHostChunkMax * (qpb::tpq::kPrepSize * sizeof(double) + sizeof(double) + sizeof(uint32_t))));
  }
  // Stages of records are uploaded, solved and downloaded on a ring of streams, so the copy engines and the SMs overlap.
  // One-launch kernels: stage sizes halve towards the end of the batch so the last kernel + download (the part that
  // cannot overlap an upload) is short.  Range-space path (three launches per stage): about eight equal stages.
  const int64_t chunk = h->host_chunk, min_chunk = 1024;
  int64_t even = ((n + h->host_stages - 1) / h->host_stages + 1023) / 1024 * 1024;
  if (even < 4096) even = 4096;
  if (even > kHostChunkMax) even = kHostChunkMax;
  int rc = QPB_SUCCESS;
  cudaError_t ce = cudaSuccess;
  for (int64_t lo = 0, m = 0; lo < n && rc == QPB_SUCCESS && ce == cudaSuccess; lo += m) {
    const int64_t left = n - lo;
    if (range_space) {
      m = left < even ? left : even;
    } else {
      m = left / 2 > chunk ? chunk : (left / 2 > min_chunk ? left / 2 : (left < min_chunk * 2 ? left : min_chunk));
      if (m > chunk) m = chunk;
    }
    const int slot = (int)(h->host_slot++ % kHostSlots);
    cudaStream_t st = h->streams[slot % h->host_cstreams], up = h->up_stream, down = h->down_stream;
    qpb_handle::TraceStage tr{ {}, slot, m };
    if (h->trace) {
      for (int k = 0; k < 4; k++) cudaEventCreate(&tr.ev[k]);
      cudaEventRecord(tr.ev[0], up);
    }
    // upload, once the slot's input buffers have been read by their last user (a wait on an event that was never
    // recorded returns at once)
    ce = cudaStreamWaitEvent(up, h->ev_in_free[slot], 0);
    if (ce == cudaSuccess && !wire) ce = cudaStreamWaitEvent(up, h->ev_solved[slot], 0);  // d_in is what the kernels read
    if (ce != cudaSuccess) break;
    if (wire)
      ce = cudaMemcpyAsync(h->d_win[slot], w_in + lo, m * sizeof(qpb_wire_state), cudaMemcpyHostToDevice, up);
    else
      ce = cudaMemcpyAsync(h->d_in[slot], h_states + lo, m * sizeof(qpb_state_rec), cudaMemcpyHostToDevice, up);
    if (ce == cudaSuccess && h_swing)
      ce = cudaMemcpyAsync(h->d_sw[slot], h_swing + lo, m * sizeof(qpb_swing_rec), cudaMemcpyHostToDevice, up);
    if (ce == cudaSuccess) ce = cudaEventRecord(h->ev_up[slot], up);
    if (ce != cudaSuccess) break;
    if (h->trace) cudaEventRecord(tr.ev[1], up);
    // kernels, once the records are there and the slot's previous results have left
    ce = cudaStreamWaitEvent(st, h->ev_up[slot], 0);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, h->ev_down[slot], 0);
    if (ce != cudaSuccess) break;
    if (wire) {
      qpb::wire_unpack_kernel<<<(unsigned)((m * 64 + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const uint64_t*>(h->d_win[slot]), reinterpret_cast<uint64_t*>(h->d_in[slot]), m);
      h->launches.fetch_add(1, std::memory_order_relaxed);
      ce = cudaGetLastError();
      if (ce == cudaSuccess && !h_swing) ce = cudaEventRecord(h->ev_in_free[slot], st);  // the wire records have been read
      if (ce != cudaSuccess) break;
    }
    qpb::PackedIO io{ h->d_in[slot], h->d_out[slot] };
    rc = launch_balance(h, io, m, h->ctas_per_sm_packed, st, path, nullptr, 0u, dn, range_space ? h->d_scratch[slot] : nullptr);
    if (rc == QPB_SUCCESS && h_swing) rc = launch_swing(h, m, h->d_in[slot], h->d_sw[slot], h->d_out[slot], st);
    if (rc != QPB_SUCCESS) break;
    if (!wire || h_swing) ce = cudaEventRecord(h->ev_in_free[slot], st);
    if (ce != cudaSuccess) break;
    if (wire) {
      qpb::wire_pack_kernel<<<(unsigned)((m * qpb::kWireOutWords + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const uint64_t*>(h->d_out[slot]), reinterpret_cast<uint64_t*>(h->d_wout[slot]), m);
      h->launches.fetch_add(1, std::memory_order_relaxed);
      ce = cudaGetLastError();
      if (ce != cudaSuccess) break;
    }
    ce = cudaEventRecord(h->ev_solved[slot], st);
    if (ce != cudaSuccess) break;
    if (h->trace) cudaEventRecord(tr.ev[2], st);
    // download
    ce = cudaStreamWaitEvent(down, h->ev_solved[slot], 0);
    if (ce != cudaSuccess) break;
    if (wire)
      ce = cudaMemcpyAsync(w_out + lo, h->d_wout[slot], m * sizeof(qpb_wire_out), cudaMemcpyDeviceToHost, down);
    else
      ce = cudaMemcpyAsync(h_out + lo, h->d_out[slot], m * sizeof(qpb_out_rec), cudaMemcpyDeviceToHost, down);
    if (ce == cudaSuccess) ce = cudaEventRecord(h->ev_down[slot], down);
    if (h->trace) {
      cudaEventRecord(tr.ev[3], down);
      h->trace_stages.push_back(tr);
    }
  }
  if (rc != QPB_SUCCESS || ce != cudaSuccess) {
    // earlier stages may still be copying into the caller's buffers: never return while they are in flight
    const std::string msg = ce != cudaSuccess ? std::string("host pipeline: ") + cudaGetErrorString(ce) : g_last_error;
    (void)sync_pipeline(h);
    return fail(ce != cudaSuccess ? QPB_ERR_CUDA : rc, msg);
  }
  return async ? QPB_SUCCESS : sync_pipeline(h);
}

}  // namespace

extern "C" {

int qpb_version(void) { return QPB_VERSION; }

#ifdef QPB_TPQ_TIMELINE
// developer build only (not declared in qpb200.h): per-CTA start / end stamps of the last loop and finishing passes
int qpb_debug_timeline(unsigned long long* out /* [4][8192] */) {
  QPB_CUDA(cudaDeviceSynchronize());
  QPB_CUDA(cudaMemcpyFromSymbol(out, qpb::tpq::g_tl, sizeof(unsigned long long) * 4 * 8192));
  static unsigned long long zero[4 * 8192];
  QPB_CUDA(cudaMemcpyToSymbol(qpb::tpq::g_tl, zero, sizeof(zero)));
  return QPB_SUCCESS;
}
#endif

const char* qpb_last_error(void) { return g_last_error.c_str(); }

int qpb_default_params(qpb_params* p) {
  if (!p) return fail(QPB_ERR_INVALID_ARG, "qpb_default_params: null pointer");
  std::memset(p, 0, sizeof(*p));
  p->mu = 0.8;
  p->mass = 11.0;
  p->fzmin = 10.0;
  p->fzmax = 120.0;
  p->Ib[0] = 0.011253;
  p->Ib[4] = 0.036203;
  p->Ib[8] = 0.042673;
  const double sd[6] = { 1.0, 1.0, 1.0, 10.0, 10.0, 5.0 };
  for (int i = 0; i < 6; i++) p->S[7 * i] = sd[i];
  for (int i = 0; i < 12; i++) p->W[13 * i] = 1e-5;
  p->kff[2] = 0.15;
  for (int i = 0; i < 3; i++) {
    p->kp_p[i] = 100.0;
    p->kd_p[i] = 50.0;
    p->kp_w[i] = 5000.0;
    p->kd_w[i] = 500.0;
  }
  const double sx[4] = { -1.0, 1.0, -1.0, 1.0 }, sy[4] = { 1.0, 1.0, -1.0, -1.0 };  // RL FL RR FR
  for (int leg = 0; leg < 4; leg++) {
    p->hip_offset[3 * leg] = sx[leg] * 0.196;
    p->hip_offset[3 * leg + 1] = sy[leg] * 0.050;
    p->hip_offset[3 * leg + 2] = 0.0;
    p->link[3 * leg] = sy[leg] * 0.077;
    p->link[3 * leg + 1] = -0.211;
    p->link[3 * leg + 2] = -0.230;
  }
  p->tau_min = -20.0;
  p->tau_max = 20.0;
  p->clamp_tau = 0;
  p->max_iter = 200;
  return QPB_SUCCESS;
}

int qpb_create(const qpb_params* params, int device, qpb_handle** out) {
  if (!params || !out) return fail(QPB_ERR_INVALID_ARG, "qpb_create: null pointer");
  *out = nullptr;
  const double* pd = reinterpret_cast<const double*>(params);
  const size_t nd = offsetof(qpb_params, clamp_tau) / sizeof(double);
  for (size_t i = 0; i < nd; i++)
    if (!std::isfinite(pd[i])) return fail(QPB_ERR_BAD_PARAMS, "qpb_create: non-finite parameter");
  if (!(params->mu > 0.0)) return fail(QPB_ERR_BAD_PARAMS, "qpb_create: mu must be > 0");
  if (!(params->fzmin <= params->fzmax)) return fail(QPB_ERR_BAD_PARAMS, "qpb_create: fzmin > fzmax");
  if (!(params->fzmax >= 0.0)) return fail(QPB_ERR_BAD_PARAMS, "qpb_create: fzmax < 0 leaves the friction pyramid empty");
  if (!(2.0 * params->mu * params->fzmax <= 1.0e6))
    return fail(QPB_ERR_BAD_PARAMS, "qpb_create: 2*mu*fzmax > 1e6: the reference's far bounds could become active");
  if (params->max_iter < 1) return fail(QPB_ERR_BAD_PARAMS, "qpb_create: max_iter < 1");
  if (!is_sympd(params->S, 6)) return fail(QPB_ERR_BAD_PARAMS, "qpb_create: S is not symmetric positive definite");
  if (!is_sympd(params->W, 12)) return fail(QPB_ERR_BAD_PARAMS, "qpb_create: W is not symmetric positive definite");

  int ndev = 0;
  QPB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(QPB_ERR_INVALID_ARG, "qpb_create: no such CUDA device");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "qpb_create: cudaSetDevice failed");
  qpb_handle* h = new (std::nothrow) qpb_handle;
  if (!h) return fail(QPB_ERR_NO_MEMORY, "qpb_create: out of host memory");
  h->device = device;
  h->params = *params;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_params, sizeof(qpb_params));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_params, params, sizeof(qpb_params), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_gains, sizeof(qpb_joint_gains));
  if (e == cudaSuccess) {
    const qpb_joint_gains g = { { 0.0, 0.0, 0.0 }, { 40.0, 40.0, 50.0 }, { 1.0, 1.0, 1.0 } };  // mit_cheetah_config.yaml:50-53
    e = cudaMemcpy(h->d_gains, &g, sizeof(g), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->d_plan, sizeof(qpb_plan_params));
  if (e == cudaSuccess) {
    qpb_plan_params pp;
    qpb_default_plan_params(&pp);
    e = cudaMemcpy(h->d_plan, &pp, sizeof(pp), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->d_tickets, 4 * kTicketSlots * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->d_tickets, 0, 4 * kTicketSlots * sizeof(unsigned long long));
  if (e == cudaSuccess)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->ctas_per_sm_packed, qpb::balance_qp_kernel<qpb::PackedIO>,
                                                      qpb::WARPS_PER_CTA * 32, 0);
  if (e == cudaSuccess)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->ctas_per_sm_split, qpb::balance_qp_kernel<qpb::SplitIO>,
                                                      qpb::WARPS_PER_CTA * 32, 0);
  if (e == cudaSuccess) {
    int a = 0, b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, qpb::balance_qp_kernel16<qpb::PackedIO>, qpb::WARPS_PER_CTA * 32, 0);
    if (e == cudaSuccess)
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, qpb::balance_qp_kernel16<qpb::SplitIO>, qpb::WARPS_PER_CTA * 32, 0);
    h->ctas_per_sm_16 = a < b ? a : b;
  }
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->ctas_per_sm_tpq[1], qpb::tpq::tpq_loop_kernel<1>, qpb::tpq::LoopShape<1>::THREADS, 0);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->ctas_per_sm_tpq[2], qpb::tpq::tpq_loop_kernel<2>, qpb::tpq::LoopShape<2>::THREADS, 0);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->ctas_per_sm_tpq[4], qpb::tpq::tpq_loop_kernel<4>, qpb::tpq::LoopShape<4>::THREADS, 0);
  if (e == cudaSuccess && qpb::tpq::StageIn<qpb::PackedIO>::on) {
    e = cudaFuncSetAttribute(qpb::tpq::tpq_setup_kernel<qpb::PackedIO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qpb::tpq::kSetupStageBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(qpb::tpq::tpq_setup_kernel<qpb::PackedIO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qpb::tpq::kSetupStageBytes);
  }
  if (e == cudaSuccess) {  // keep freed scratch in the stream-ordered pool instead of returning it to the OS after every call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
  }
  if (e != cudaSuccess || h->ctas_per_sm_packed < 1 || h->ctas_per_sm_split < 1 || h->ctas_per_sm_16 < 1 ||
      h->ctas_per_sm_tpq[1] < 1 || h->ctas_per_sm_tpq[2] < 1 || h->ctas_per_sm_tpq[4] < 1) {
    const std::string msg = std::string("qpb_create: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit");
    if (h->d_params) cudaFree(h->d_params);
    if (h->d_tickets) cudaFree(h->d_tickets);
    if (h->d_gains) cudaFree(h->d_gains);
    if (h->d_plan) cudaFree(h->d_plan);
    delete h;
    return fail(QPB_ERR_CUDA, msg);
  }
  h->num_sms = prop.multiProcessorCount;
  // The thread-per-QP kernel needs W = w I (range-space form) and fzmin >= 0 (its working sets stay within the faces
  // of the truncated pyramid, qpb_tpq_core.h); everything else takes the general half-warp kernel.
  const bool fast_ok = qpb::tpq::make_fast_params(*params, h->fast) && params->fzmin >= 0.0;
  h->edge = qpb::tpq::make_edge_params(*params);
  h->qps_per_warp = fast_ok ? 32 : 2;
  if (const char* env = std::getenv("QPB_QPS_PER_WARP")) {
    const int v = std::atoi(env);
    if (v == 1 || v == 2 || (v == 32 && fast_ok)) h->qps_per_warp = v;
  }
  if (const char* env = std::getenv("QPB_TPQ_LPQ")) {
    const int v = std::atoi(env);
    if (v == 1 || v == 2 || v == 4) h->tpq_lpq = v;
  }
  if (const char* env = std::getenv("QPB_TPQ_PDL")) h->tpq_pdl = std::atoi(env) != 0;
  if (const char* env = std::getenv("QPB_TPQ_MIN_N")) {
    const long long v = std::atoll(env);
    if (v >= 0) h->tpq_min_n = v;
  }
  if (const char* env = std::getenv("QPB_TPQ_WARM_DEFER_MIN")) {
    const long long v = std::atoll(env);
    if (v >= 0) h->tpq_warm_defer_min = v;
  }
  if (const char* env = std::getenv("QPB_TPQ_ONE_MAX")) {
    const long long v = std::atoll(env);
    if (v >= 0) h->tpq_one_max = v;
  }
  if (const char* env = std::getenv("QPB_ZEROCOPY")) h->zero_copy = std::atoi(env);
  if (const char* env = std::getenv("QPB_SMALL_POLL")) h->small_poll = std::atoi(env);
  if (const char* env = std::getenv("QPB_HOST_TRACE")) h->trace = std::atoi(env);
  if (const char* env = std::getenv("QPB_HOST_CSTREAMS")) {
    const int v = std::atoi(env);
    if (v >= 1 && v <= kHostSlots) h->host_cstreams = v;
  }
  if (const char* env = std::getenv("QPB_HOST_STAGES")) {
    const int v = std::atoi(env);
    if (v >= 1 && v <= 64) h->host_stages = v;
  }
  if (const char* env = std::getenv("QPB_HOST_CHUNK")) {
    const long long v = std::atoll(env);
    if (v >= 256 && v <= kHostChunkMax) h->host_chunk = v;
  }
  *out = h;
  return QPB_SUCCESS;
}

int qpb_destroy(qpb_handle* h) {
  if (!h) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  for (int s = 0; s < kHostSlots; s++) {
    if (h->streams[s]) cudaStreamDestroy(h->streams[s]);
    for (cudaEvent_t e : { h->ev_up[s], h->ev_in_free[s], h->ev_solved[s], h->ev_down[s] })
      if (e) cudaEventDestroy(e);
    if (h->d_in[s]) cudaFree(h->d_in[s]);
    if (h->d_out[s]) cudaFree(h->d_out[s]);
    if (h->d_sw[s]) cudaFree(h->d_sw[s]);
    if (h->d_scratch[s]) cudaFree(h->d_scratch[s]);
    if (h->d_win[s]) cudaFree(h->d_win[s]);
    if (h->d_wout[s]) cudaFree(h->d_wout[s]);
  }
  if (h->up_stream) cudaStreamDestroy(h->up_stream);
  if (h->down_stream) cudaStreamDestroy(h->down_stream);
  if (h->d_params) cudaFree(h->d_params);
  if (h->d_tickets) cudaFree(h->d_tickets);
  if (h->d_gains) cudaFree(h->d_gains);
  if (h->d_plan) cudaFree(h->d_plan);
  if (h->h_small) cudaFreeHost(h->h_small);
  delete h;
  return QPB_SUCCESS;
}

int qpb_control_batch_packed(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, qpb_out_rec* d_out,
                             void* stream) {
  if (!h || n < 0 || (n > 0 && (!d_states || !d_out)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch_packed: bad argument");
  if ((reinterpret_cast<uintptr_t>(d_states) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
    return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch_packed: records must be 16-byte aligned");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  qpb::PackedIO io{ d_states, d_out };
  return launch_balance(h, io, n, h->ctas_per_sm_packed, static_cast<cudaStream_t>(stream));
}

int qpb_control_batch(qpb_handle* h, int64_t n, const double* Rwb, const double* Rwb_d, const double* x,
                      const double* xdot, const double* w, const double* x_d, const double* xdot_d,
                      const double* w_d, const double* feet_body, const uint8_t* contact, const double* q,
                      double* grf_body, double* tau, int32_t* status, void* stream) {
  if (!h || n < 0) return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch: bad argument");
  if (n > 0 && (!Rwb || !Rwb_d || !x || !xdot || !w || !x_d || !xdot_d || !w_d || !feet_body || !contact || !q ||
                !grf_body))
    return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch: null array");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  qpb::SplitIO io{ Rwb, Rwb_d, x, xdot, w, x_d, xdot_d, w_d, feet_body, q, contact, grf_body, tau, status };
  return launch_balance(h, io, n, h->ctas_per_sm_split, static_cast<cudaStream_t>(stream));
}

int qpb_control_batch_host(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, qpb_out_rec* h_out) {
  if (!h || n < 0 || (n > 0 && (!h_states || !h_out)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch_host: bad argument");
  return host_pipeline(h, n, h_states, nullptr, h_out);
}

int qpb_internal_host_shard(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, const qpb_swing_rec* h_swing,
                            qpb_out_rec* h_out, int64_t whole_n, int whole_warm) {
  if (!h || n < 0 || (n > 0 && (!h_states || !h_out))) return fail(QPB_ERR_INVALID_ARG, "host shard: bad argument");
  return host_pipeline(h, n, h_states, h_swing, h_out, false, whole_n, whole_warm);
}

int qpb_control_batch_host_async(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, qpb_out_rec* h_out) {
  if (!h || n < 0 || (n > 0 && (!h_states || !h_out)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch_host_async: bad argument");
  return host_pipeline(h, n, h_states, nullptr, h_out, true);
}

int qpb_control_batch_wire_host(qpb_handle* h, int64_t n, const qpb_wire_state* h_states, qpb_wire_out* h_out) {
  if (!h || n < 0 || (n > 0 && (!h_states || !h_out)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch_wire_host: bad argument");
  return host_pipeline(h, n, nullptr, nullptr, nullptr, false, -1, -1, h_states, h_out);
}

int qpb_control_batch_wire_host_async(qpb_handle* h, int64_t n, const qpb_wire_state* h_states, qpb_wire_out* h_out) {
  if (!h || n < 0 || (n > 0 && (!h_states || !h_out)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_control_batch_wire_host_async: bad argument");
  return host_pipeline(h, n, nullptr, nullptr, nullptr, true, -1, -1, h_states, h_out);
}

int qpb_host_sync(qpb_handle* h) {
  if (!h) return fail(QPB_ERR_INVALID_ARG, "qpb_host_sync: null handle");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  return sync_pipeline(h);
}

int qpb_set_warm_batches(qpb_handle* h, int on) {
  if (!h) return fail(QPB_ERR_INVALID_ARG, "qpb_set_warm_batches: null handle");
  h->warm_batches = on ? 1 : 0;
  return QPB_SUCCESS;
}

int qpb_set_joint_gains(qpb_handle* h, const qpb_joint_gains* gains) {
  if (!h || !gains) return fail(QPB_ERR_INVALID_ARG, "qpb_set_joint_gains: null pointer");
  const double* g = reinterpret_cast<const double*>(gains);
  for (int i = 0; i < 9; i++)
    if (!std::isfinite(g[i])) return fail(QPB_ERR_BAD_PARAMS, "qpb_set_joint_gains: non-finite gain");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  QPB_CUDA(cudaDeviceSynchronize());  // earlier launches may still read the old gains
  QPB_CUDA(cudaMemcpy(h->d_gains, gains, sizeof(qpb_joint_gains), cudaMemcpyHostToDevice));
  return QPB_SUCCESS;
}

int qpb_tick_batch_packed(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, const qpb_swing_rec* d_swing,
                          qpb_out_rec* d_out, void* stream) {
  if (!h || n < 0 || (n > 0 && (!d_states || !d_swing || !d_out)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_tick_batch_packed: bad argument");
  if ((reinterpret_cast<uintptr_t>(d_states) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
    return fail(QPB_ERR_INVALID_ARG, "qpb_tick_batch_packed: records must be 16-byte aligned");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  qpb::PackedIO io{ d_states, d_out };
  const int rc = launch_balance(h, io, n, h->ctas_per_sm_packed, static_cast<cudaStream_t>(stream));
  if (rc != QPB_SUCCESS) return rc;
  return launch_swing(h, n, d_states, d_swing, d_out, static_cast<cudaStream_t>(stream));
}

int qpb_tick_batch_host(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, const qpb_swing_rec* h_swing,
                        qpb_out_rec* h_out) {
  if (!h || n < 0 || (n > 0 && (!h_states || !h_swing || !h_out)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_tick_batch_host: bad argument");
  return host_pipeline(h, n, h_states, h_swing, h_out);
}

int qpb_jt_batch(qpb_handle* h, int64_t n, const double* q, const double* grf_body, const uint8_t* contact,
                 double* tau, void* stream) {
  if (!h || n < 0 || (n > 0 && (!q || !grf_body || !tau))) return fail(QPB_ERR_INVALID_ARG, "qpb_jt_batch: bad argument");
  if (n == 0) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  const int64_t nlegs = 4 * n;
  const int threads = 128;
  const int64_t blocks = (nlegs + threads - 1) / threads;
  qpb::jt_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->d_params, q, grf_body, contact,
                                                                                     tau, nlegs);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  QPB_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

int qpb_fk_batch(qpb_handle* h, int64_t n, const double* q, double* feet_body, void* stream) {
  if (!h || n < 0 || (n > 0 && (!q || !feet_body))) return fail(QPB_ERR_INVALID_ARG, "qpb_fk_batch: bad argument");
  if (n == 0) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  const int64_t nlegs = 4 * n;
  const int threads = 128;
  const int64_t blocks = (nlegs + threads - 1) / threads;
  qpb::fk_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->d_params, q, feet_body, nlegs);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  QPB_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

int qpb_jt_batch_host(qpb_handle* h, int64_t n, const double* q, const double* grf_body, const uint8_t* contact,
                      double* tau) {
  if (!h || n < 0 || (n > 0 && (!q || !grf_body || !tau))) return fail(QPB_ERR_INVALID_ARG, "qpb_jt_batch_host: bad argument");
  if (n == 0) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  const size_t nb = (size_t)n * 12 * sizeof(double);
  if (h->zero_copy && n <= kSmallCall) {  // latency path: pinned, mapped staging block, no cudaMalloc / cudaMemcpy
    const int rc0 = ensure_small(h);
    if (rc0 != QPB_SUCCESS) return rc0;
    double* hs = reinterpret_cast<double*>(h->h_small);
    double* dsm = reinterpret_cast<double*>(h->d_small);
    std::memcpy(hs, q, nb);
    std::memcpy(hs + (size_t)n * 12, grf_body, nb);
    uint8_t* hc = reinterpret_cast<uint8_t*>(hs + 3 * (size_t)n * 12);
    if (contact) std::memcpy(hc, contact, (size_t)n * 4);
    const int rc = qpb_jt_batch(h, n, dsm, dsm + (size_t)n * 12, contact ? reinterpret_cast<uint8_t*>(dsm + 3 * (size_t)n * 12) : nullptr,
                                dsm + 2 * (size_t)n * 12, h->streams[0]);
    if (rc != QPB_SUCCESS) return rc;
    QPB_CUDA(cudaStreamSynchronize(h->streams[0]));
    std::memcpy(tau, hs + 2 * (size_t)n * 12, nb);
    return QPB_SUCCESS;
  }
  double* d = nullptr;  // [q | f | tau | contact]
  QPB_CUDA(cudaMalloc(&d, 3 * nb + (size_t)n * 4));
  uint8_t* dc = reinterpret_cast<uint8_t*>(d + 3 * (size_t)n * 12);
  cudaError_t e = cudaMemcpy(d, q, nb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + (size_t)n * 12, grf_body, nb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && contact) e = cudaMemcpy(dc, contact, (size_t)n * 4, cudaMemcpyHostToDevice);
  int rc = QPB_SUCCESS;
  if (e == cudaSuccess) rc = qpb_jt_batch(h, n, d, d + (size_t)n * 12, contact ? dc : nullptr, d + 2 * (size_t)n * 12, nullptr);
  if (e == cudaSuccess && rc == QPB_SUCCESS) e = cudaMemcpy(tau, d + 2 * (size_t)n * 12, nb, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(QPB_ERR_CUDA, std::string("qpb_jt_batch_host: ") + cudaGetErrorString(e));
  return rc;
}

int qpb_fk_batch_host(qpb_handle* h, int64_t n, const double* q, double* feet_body) {
  if (!h || n < 0 || (n > 0 && (!q || !feet_body))) return fail(QPB_ERR_INVALID_ARG, "qpb_fk_batch_host: bad argument");
  if (n == 0) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  const size_t nb = (size_t)n * 12 * sizeof(double);
  if (h->zero_copy && n <= kSmallCall) {  // latency path, as in qpb_jt_batch_host
    const int rc0 = ensure_small(h);
    if (rc0 != QPB_SUCCESS) return rc0;
    double* hs = reinterpret_cast<double*>(h->h_small);
    double* dsm = reinterpret_cast<double*>(h->d_small);
    std::memcpy(hs, q, nb);
    const int rc = qpb_fk_batch(h, n, dsm, dsm + (size_t)n * 12, h->streams[0]);
    if (rc != QPB_SUCCESS) return rc;
    QPB_CUDA(cudaStreamSynchronize(h->streams[0]));
    std::memcpy(feet_body, hs + (size_t)n * 12, nb);
    return QPB_SUCCESS;
  }
  double* d = nullptr;
  QPB_CUDA(cudaMalloc(&d, 2 * nb));
  cudaError_t e = cudaMemcpy(d, q, nb, cudaMemcpyHostToDevice);
  int rc = QPB_SUCCESS;
  if (e == cudaSuccess) rc = qpb_fk_batch(h, n, d, d + (size_t)n * 12, nullptr);
  if (e == cudaSuccess && rc == QPB_SUCCESS) e = cudaMemcpy(feet_body, d + (size_t)n * 12, nb, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(QPB_ERR_CUDA, std::string("qpb_fk_batch_host: ") + cudaGetErrorString(e));
  return rc;
}

int qpb_default_plan_params(qpb_plan_params* p) {
  if (!p) return fail(QPB_ERR_INVALID_ARG, "qpb_default_plan_params: null pointer");
  std::memset(p, 0, sizeof(*p));
  p->k_raibert = 0.01;
  p->g = 9.81;
  const double sx[4] = { -1.0, 1.0, -1.0, 1.0 }, sy[4] = { 1.0, 1.0, -1.0, -1.0 };  // RL FL RR FR
  for (int leg = 0; leg < 4; leg++) {
    p->thigh_offset[3 * leg] = sx[leg] * 0.196;
    p->thigh_offset[3 * leg + 1] = sy[leg] * 0.127;
    p->thigh_offset[3 * leg + 2] = 0.0;
  }
  p->height = 0.08;
  p->t_swing = 0.18;
  p->t_stance = 0.8;
  return QPB_SUCCESS;
}

int qpb_set_plan_params(qpb_handle* h, const qpb_plan_params* p) {
  if (!h || !p) return fail(QPB_ERR_INVALID_ARG, "qpb_set_plan_params: null pointer");
  const double* v = reinterpret_cast<const double*>(p);
  for (size_t i = 0; i < sizeof(*p) / sizeof(double); i++)
    if (!std::isfinite(v[i])) return fail(QPB_ERR_BAD_PARAMS, "qpb_set_plan_params: non-finite parameter");
  if (!(p->g > 0.0) || !(p->t_swing > 0.0) || !(p->t_stance >= 0.0))
    return fail(QPB_ERR_BAD_PARAMS, "qpb_set_plan_params: g and t_swing must be > 0, t_stance >= 0");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  QPB_CUDA(cudaDeviceSynchronize());  // earlier launches may still read the old constants
  QPB_CUDA(cudaMemcpy(h->d_plan, p, sizeof(*p), cudaMemcpyHostToDevice));
  return QPB_SUCCESS;
}

int qpb_plan_batch(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, qpb_plan_rec* d_plan, qpb_swing_rec* d_swing,
                   void* stream) {
  if (!h || n < 0 || (n > 0 && (!d_states || !d_plan || !d_swing))) return fail(QPB_ERR_INVALID_ARG, "qpb_plan_batch: bad argument");
  if (n == 0) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  const int threads = 128;
  qpb::plan_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_plan, d_states, d_plan, d_swing, n);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  QPB_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

int qpb_adapt_inputs_batch(qpb_handle* h, int64_t n, const qpb_com_msg* d_com, const qpb_joint_msg* d_joints,
                           qpb_state_rec* d_states, qpb_swing_rec* d_swing, void* stream) {
  if (!h || n < 0 || (n > 0 && (!d_com || !d_joints || !d_states || !d_swing)))
    return fail(QPB_ERR_INVALID_ARG, "qpb_adapt_inputs_batch: bad argument");
  if (n == 0) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  const int threads = qpb::kXformThreads;
  const unsigned grid = (unsigned)((n + threads - 1) / threads);
  // 32-byte-aligned arrays (any cudaMalloc'd buffer): 256-bit loads/stores; otherwise the 16-byte ABI minimum
  const bool v256 = ((reinterpret_cast<uintptr_t>(d_com) | reinterpret_cast<uintptr_t>(d_joints) |
                      reinterpret_cast<uintptr_t>(d_states) | reinterpret_cast<uintptr_t>(d_swing)) & 31u) == 0;
  if ((reinterpret_cast<uintptr_t>(d_states) | reinterpret_cast<uintptr_t>(d_swing) | reinterpret_cast<uintptr_t>(d_joints)) & 15u)
    return fail(QPB_ERR_INVALID_ARG, "qpb_adapt_inputs_batch: records must be 16-byte aligned");
  if (v256)
    qpb::adapt_kernel<true><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->d_params, d_com, d_joints, d_states, d_swing, n);
  else
    qpb::adapt_kernel<false><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->d_params, d_com, d_joints, d_states, d_swing, n);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  QPB_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

int qpb_torque_cmd_batch(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, const qpb_out_rec* d_out,
                         qpb_torque_cmd* d_cmd, void* stream) {
  if (!h || n < 0 || (n > 0 && (!d_states || !d_out || !d_cmd))) return fail(QPB_ERR_INVALID_ARG, "qpb_torque_cmd_batch: bad argument");
  if (n == 0) return QPB_SUCCESS;
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  const int threads = qpb::kXformThreads;
  const unsigned grid = (unsigned)((n + threads - 1) / threads);
  if ((reinterpret_cast<uintptr_t>(d_states) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_cmd)) & 15u)
    return fail(QPB_ERR_INVALID_ARG, "qpb_torque_cmd_batch: records must be 16-byte aligned");
  const bool v256 = ((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_cmd)) & 31u) == 0;
  if (v256)
    qpb::torque_cmd_kernel<true><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->d_params, d_states, d_out, d_cmd, n);
  else
    qpb::torque_cmd_kernel<false><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->d_params, d_states, d_out, d_cmd, n);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  QPB_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

int qpb_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(QPB_ERR_INVALID_ARG, "qpb_host_alloc: null pointer");
  QPB_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));  // visible to every device (qpb_multi_*)
  return QPB_SUCCESS;
}

int qpb_host_free(void* ptr) {
  if (!ptr) return QPB_SUCCESS;
  QPB_CUDA(cudaFreeHost(ptr));
  return QPB_SUCCESS;
}

int64_t qpb_launch_count(const qpb_handle* h) { return h ? h->launches.load(std::memory_order_relaxed) : 0; }

}  // extern "C"
