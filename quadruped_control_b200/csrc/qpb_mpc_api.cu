// qpb_mpc_api.cu -- C ABI of the 10-step convex-MPC QP (include/qpb200.h, qpb_mpc_*) on top of qpb_mpc.cuh.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "qpb_internal.h"
#include "qpb_mpc.cuh"

namespace {

#define MPC_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return qpb_internal_fail(QPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

constexpr int kSlots = 3;            // streams / staging buffers of the host-buffer pipeline
constexpr int64_t kChunk = 4096;     // records per stage (8.9 MB in, 4.2 MB out)
constexpr uint32_t kTicketSlots = 1024;

struct Guard {
  int prev = -1;
  bool ok = true;
  explicit Guard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~Guard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

bool inv3(const double* M, double* out) {
  const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  if (!(std::fabs(det) > 0.0) || !std::isfinite(1.0 / det)) return false;
  out[0] = (e * i - f * h) / det; out[1] = (c * h - b * i) / det; out[2] = (b * f - c * e) / det;
  out[3] = (f * g - d * i) / det; out[4] = (a * i - c * g) / det; out[5] = (c * d - a * f) / det;
  out[6] = (d * h - e * g) / det; out[7] = (b * g - a * h) / det; out[8] = (a * e - b * d) / det;
  return true;
}

}  // namespace

struct qpb_mpc_handle {
  int device = 0;
  int num_sms = 0;
  qpbmpc::DevParams dp;
  unsigned long long* d_tickets = nullptr;
  std::atomic<uint32_t> ticket_slot{ 0 };
  std::atomic<int64_t> launches{ 0 };
  cudaStream_t streams[kSlots] = {};
  qpb_mpc_rec* d_in[kSlots] = {};
  qpb_mpc_out_rec* d_out[kSlots] = {};
};

namespace {

int launch_mpc(qpb_mpc_handle* h, int64_t n, const qpb_mpc_rec* d_recs, qpb_mpc_out_rec* d_out, cudaStream_t stream) {
  if (n == 0) return QPB_SUCCESS;
  if (n > ((int64_t)1 << 31) - 4096) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "more than 2^31 records in one call: split the batch");
  const int grid = (int)(n < h->num_sms ? n : h->num_sms);  // one persistent CTA per SM (215 KB of shared memory each)
  const uint32_t slot = h->ticket_slot.fetch_add(1, std::memory_order_relaxed) % kTicketSlots;
  unsigned long long* t0 = h->d_tickets + 2 * (size_t)slot;  // {work counter, CTAs finished}; the kernel re-arms both
  qpbmpc::mpc_qp_kernel<256><<<grid, 256, sizeof(qpbmpc::Smem), stream>>>(h->dp, d_recs, d_out, n, t0);
  h->launches.fetch_add(1, std::memory_order_relaxed);
  MPC_CUDA(cudaGetLastError());
  return QPB_SUCCESS;
}

}  // namespace

extern "C" {

int qpb_mpc_default_params(qpb_mpc_params* p) {
  if (!p) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_mpc_default_params: null pointer");
  std::memset(p, 0, sizeof(*p));
  p->mu = 0.6;
  p->mass = 11.0;
  p->fzmin = 10.0;
  p->fzmax = 120.0;
  p->Ib[0] = 0.011253;
  p->Ib[4] = 0.036203;
  p->Ib[8] = 0.042673;
  p->dt = 0.03;
  const double Lw[13] = { 0.25, 0.25, 10.0, 2.0, 2.0, 50.0, 0.0, 0.0, 0.3, 0.2, 0.2, 0.1, 0.0 };
  std::memcpy(p->Lw, Lw, sizeof(Lw));
  p->alpha = 4e-5;
  p->max_iter = 1000;
  return QPB_SUCCESS;
}

int qpb_mpc_create(const qpb_mpc_params* params, int device, qpb_mpc_handle** out) {
  if (!params || !out) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_mpc_create: null pointer");
  *out = nullptr;
  const double* pd = reinterpret_cast<const double*>(params);
  for (size_t i = 0; i < offsetof(qpb_mpc_params, max_iter) / sizeof(double); i++)
    if (!std::isfinite(pd[i])) return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: non-finite parameter");
  if (!(params->mu > 0.0)) return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: mu must be > 0");
  if (!(params->fzmin <= params->fzmax)) return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: fzmin > fzmax");
  if (!(params->fzmax >= 0.0)) return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: fzmax < 0 leaves the friction pyramid empty");
  if (!(2.0 * params->mu * params->fzmax <= 1.0e6))
    return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: 2*mu*fzmax > 1e6: the reference's far bounds could become active");
  if (!(params->mass > 0.0) || !(params->dt > 0.0) || !(params->alpha > 0.0))
    return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: mass, dt and alpha must be > 0");
  for (int i = 0; i < 12; i++)
    if (!(params->Lw[i] >= 0.0)) return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: negative state weight");
  if (params->max_iter < 1) return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: max_iter < 1");
  qpbmpc::DevParams dp;
  std::memset(&dp, 0, sizeof(dp));
  if (!inv3(params->Ib, dp.Ibinv)) return qpb_internal_fail(QPB_ERR_BAD_PARAMS, "qpb_mpc_create: Ib is singular");
  dp.mu = params->mu; dp.mass = params->mass; dp.fzmin = params->fzmin; dp.fzmax = params->fzmax;
  dp.dt = params->dt;
  for (int i = 0; i < 12; i++) dp.Lw[i] = params->Lw[i];
  for (int i = 0; i < 3; i++) dp.sLw[i] = std::sqrt(params->Lw[i]);
  dp.alpha = params->alpha;
  dp.max_iter = params->max_iter;

  int ndev = 0;
  MPC_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_mpc_create: no such CUDA device");
  Guard guard(device);
  if (!guard.ok) return qpb_internal_fail(QPB_ERR_CUDA, "qpb_mpc_create: cudaSetDevice failed");
  qpb_mpc_handle* h = new (std::nothrow) qpb_mpc_handle;
  if (!h) return qpb_internal_fail(QPB_ERR_NO_MEMORY, "qpb_mpc_create: out of host memory");
  h->device = device;
  h->dp = dp;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(qpbmpc::mpc_qp_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(qpbmpc::Smem));

  if (e == cudaSuccess) e = cudaMalloc(&h->d_tickets, 2 * kTicketSlots * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->d_tickets, 0, 2 * kTicketSlots * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    const std::string msg = std::string("qpb_mpc_create: ") + cudaGetErrorString(e);
    if (h->d_tickets) cudaFree(h->d_tickets);
    delete h;
    return qpb_internal_fail(QPB_ERR_CUDA, msg);
  }
  h->num_sms = prop.multiProcessorCount;
  *out = h;
  return QPB_SUCCESS;
}

int qpb_mpc_destroy(qpb_mpc_handle* h) {
  if (!h) return QPB_SUCCESS;
  Guard guard(h->device);
  for (int s = 0; s < kSlots; s++) {
    if (h->streams[s]) cudaStreamDestroy(h->streams[s]);
    if (h->d_in[s]) cudaFree(h->d_in[s]);
    if (h->d_out[s]) cudaFree(h->d_out[s]);
  }
  if (h->d_tickets) cudaFree(h->d_tickets);
  delete h;
  return QPB_SUCCESS;
}

int qpb_mpc_batch_packed(qpb_mpc_handle* h, int64_t n, const qpb_mpc_rec* d_recs, qpb_mpc_out_rec* d_out, void* stream) {
  if (!h || n < 0 || (n > 0 && (!d_recs || !d_out))) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_mpc_batch_packed: bad argument");
  if ((reinterpret_cast<uintptr_t>(d_recs) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
    return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_mpc_batch_packed: records must be 16-byte aligned");
  Guard guard(h->device);
  if (!guard.ok) return qpb_internal_fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  return launch_mpc(h, n, d_recs, d_out, static_cast<cudaStream_t>(stream));
}

int qpb_mpc_batch_host(qpb_mpc_handle* h, int64_t n, const qpb_mpc_rec* h_recs, qpb_mpc_out_rec* h_out) {
  if (!h || n < 0 || (n > 0 && (!h_recs || !h_out))) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_mpc_batch_host: bad argument");
  if (n == 0) return QPB_SUCCESS;
  Guard guard(h->device);
  if (!guard.ok) return qpb_internal_fail(QPB_ERR_CUDA, "cudaSetDevice failed");
  for (int s = 0; s < kSlots; s++) {
    if (!h->streams[s]) MPC_CUDA(cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking));
    if (!h->d_in[s]) MPC_CUDA(cudaMalloc(&h->d_in[s], kChunk * sizeof(qpb_mpc_rec)));
    if (!h->d_out[s]) MPC_CUDA(cudaMalloc(&h->d_out[s], kChunk * sizeof(qpb_mpc_out_rec)));
  }
  // upload / solve / download of successive stages overlap on a ring of streams; the solve dominates by far
  int slot = 0, rc = QPB_SUCCESS;
  cudaError_t ce = cudaSuccess;
  for (int64_t lo = 0; lo < n && rc == QPB_SUCCESS && ce == cudaSuccess; lo += kChunk, slot = (slot + 1) % kSlots) {
    const int64_t m = n - lo < kChunk ? n - lo : kChunk;
    cudaStream_t st = h->streams[slot];
    ce = cudaMemcpyAsync(h->d_in[slot], h_recs + lo, m * sizeof(qpb_mpc_rec), cudaMemcpyHostToDevice, st);
    if (ce != cudaSuccess) break;
    rc = launch_mpc(h, m, h->d_in[slot], h->d_out[slot], st);
    if (rc != QPB_SUCCESS) break;
    ce = cudaMemcpyAsync(h_out + lo, h->d_out[slot], m * sizeof(qpb_mpc_out_rec), cudaMemcpyDeviceToHost, st);
  }
  // never return while an earlier stage may still be copying out of / into the caller's buffers -- also not on an error
  for (int s = 0; s < kSlots; s++) {
    const cudaError_t es = cudaStreamSynchronize(h->streams[s]);
    if (ce == cudaSuccess) ce = es;
  }
  if (rc != QPB_SUCCESS) return rc;
  if (ce != cudaSuccess) return qpb_internal_fail(QPB_ERR_CUDA, std::string("qpb_mpc_batch_host: ") + cudaGetErrorString(ce));
  return QPB_SUCCESS;
}

#ifdef QPB_MPC_PROFILE
// developer build only: read and reset the per-phase cycle counters (not declared in qpb200.h)
int qpb_mpc_debug_profile(unsigned long long out[24]) {
  MPC_CUDA(cudaDeviceSynchronize());
  MPC_CUDA(cudaMemcpyFromSymbol(out, qpbmpc::g_mpc_prof, 24 * sizeof(unsigned long long)));
  unsigned long long zero[24] = {};
  MPC_CUDA(cudaMemcpyToSymbol(qpbmpc::g_mpc_prof, zero, sizeof(zero)));
  return QPB_SUCCESS;
}
#endif

int64_t qpb_mpc_launch_count(const qpb_mpc_handle* h) { return h ? h->launches.load(std::memory_order_relaxed) : 0; }

}  // extern "C"
