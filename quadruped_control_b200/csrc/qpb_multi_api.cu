// qpb_multi_api.cu -- single-process multi-GPU form of the host-buffer entry points.
//
// Every QP is independent, so the batch is cut into contiguous shards [lo_r, hi_r), one per device, and there is
// no data-path collective: each shard goes through that device's own qpb_handle (qpb_control_batch_host /
// qpb_tick_batch_host) on a persistent host thread bound to the device.  A C++ caller (the ROS node, a simulator
// bridge) gets the whole box behind one call, without a process launcher.  Shard boundaries are the ones
// quadruped_control_b200/sharding.py::shard_range uses for the one-process-per-GPU (torchrun) path.
#include <cuda_runtime.h>

#include <condition_variable>
#include <exception>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qpb200.h"
#include "qpb_internal.h"

struct qpb_multi_handle {
  struct Job {
    int64_t n = 0;
    const qpb_state_rec* states = nullptr;
    const qpb_swing_rec* swing = nullptr;  // non-null: whole tick
    qpb_out_rec* out = nullptr;
    int64_t whole_n = 0;  // size of the batch this shard is cut from, and whether its first record carries a warm-start
    int whole_warm = 0;   // word: every shard takes the kernels the whole batch would take on one device
  };
  struct Worker {
    qpb_handle* h = nullptr;
    int device = 0;
    std::thread thread;
    Job job;
    int rc = 0;
    std::string err;
  };
  std::vector<Worker> workers;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t epoch = 0;  // bumped once per call; a worker runs its job when it sees a new epoch
  int pending = 0;
  bool quit = false;
  std::mutex call_mu;  // one call at a time per handle
};

namespace {

void shard_range(int64_t n, int r, int g, int64_t* lo, int64_t* hi) {
  const int64_t base = n / g, rem = n % g;
  *lo = r * base + (r < rem ? r : rem);
  *hi = *lo + base + (r < rem ? 1 : 0);
}

void worker_main(qpb_multi_handle* m, int r) {
  qpb_multi_handle::Worker& w = m->workers[r];
  cudaSetDevice(w.device);
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_go.wait(lk, [&] { return m->quit || m->epoch != seen; });
      if (m->quit) return;
      seen = m->epoch;
    }
    const qpb_multi_handle::Job& j = w.job;
    int rc = QPB_SUCCESS;
    if (j.n > 0)
      rc = qpb_internal_host_shard(w.h, j.n, j.states, j.swing, j.out, j.whole_n, j.whole_warm);
    w.rc = rc;
    w.err = rc == QPB_SUCCESS ? "" : qpb_last_error();  // the error text is thread-local: carry it to the caller
    {
      std::lock_guard<std::mutex> lk(m->mu);
      if (--m->pending == 0) m->cv_done.notify_all();
    }
  }
}

int run_sharded(qpb_multi_handle* m, int64_t n, const qpb_state_rec* states, const qpb_swing_rec* swing, qpb_out_rec* out) {
  std::lock_guard<std::mutex> call(m->call_mu);
  const int g = (int)m->workers.size();
  {
    std::lock_guard<std::mutex> lk(m->mu);
    for (int r = 0; r < g; r++) {
      int64_t lo, hi;
      shard_range(n, r, g, &lo, &hi);
      qpb_multi_handle::Job& j = m->workers[r].job;
      j.n = hi - lo;
      j.states = states + lo;
      j.swing = swing ? swing + lo : nullptr;
      j.out = out + lo;
      j.whole_n = n;
      j.whole_warm = (states[0].pad[3] & 0x80u) != 0;
    }
    m->pending = g;
    m->epoch++;
  }
  m->cv_go.notify_all();
  {
    std::unique_lock<std::mutex> lk(m->mu);
    m->cv_done.wait(lk, [&] { return m->pending == 0; });
  }
  for (int r = 0; r < g; r++)
    if (m->workers[r].rc != QPB_SUCCESS)
      return qpb_internal_fail(m->workers[r].rc, "shard " + std::to_string(r) + " (device " +
                                                     std::to_string(m->workers[r].device) + "): " + m->workers[r].err);
  return QPB_SUCCESS;
}

}  // namespace

extern "C" {

int qpb_device_count(int* count) {
  if (!count) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_device_count: null argument");
  int c = 0;
  const cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *count = 0;
    return qpb_internal_fail(QPB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = c;
  return QPB_SUCCESS;
}

int qpb_multi_shard_range(int64_t n, int shard, int num_shards, int64_t* lo, int64_t* hi) {
  if (n < 0 || num_shards < 1 || shard < 0 || shard >= num_shards || !lo || !hi)
    return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_multi_shard_range: bad argument");
  shard_range(n, shard, num_shards, lo, hi);
  return QPB_SUCCESS;
}

int qpb_multi_create(const qpb_params* params, const int* devices, int num_devices, qpb_multi_handle** out) {
  if (!params || !out) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_multi_create: null argument");
  *out = nullptr;
  if (devices && num_devices < 1) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_multi_create: num_devices < 1");
  qpb_multi_handle* m = nullptr;
  // The C ABI never throws: vector growth, std::string and std::thread construction can (bad_alloc, system_error), so
  // everything that allocates sits in one try block, and a failure at any point unwinds what was already built --
  // started workers are told to quit and joined, created handles destroyed.
  try {
    std::vector<int> devs;
    if (devices) {
      devs.assign(devices, devices + num_devices);
    } else {
      int c = 0;
      const int rc = qpb_device_count(&c);
      if (rc != QPB_SUCCESS) return rc;
      if (c < 1) return qpb_internal_fail(QPB_ERR_CUDA, "qpb_multi_create: no CUDA device (there is no CPU fallback)");
      for (int d = 0; d < c; d++) devs.push_back(d);
    }
    m = new (std::nothrow) qpb_multi_handle;
    if (!m) return qpb_internal_fail(QPB_ERR_NO_MEMORY, "qpb_multi_create: out of host memory");
    m->workers.resize(devs.size());
    for (size_t r = 0; r < devs.size(); r++) {
      m->workers[r].device = devs[r];
      const int rc = qpb_create(params, devs[r], &m->workers[r].h);
      if (rc != QPB_SUCCESS) {
        const std::string msg = qpb_last_error();
        qpb_multi_destroy(m);  // destroys the handles created so far (no worker has been started yet)
        return qpb_internal_fail(rc, "qpb_multi_create: device " + std::to_string(devs[r]) + ": " + msg);
      }
    }
    for (size_t r = 0; r < devs.size(); r++) m->workers[r].thread = std::thread(worker_main, m, (int)r);
  } catch (const std::exception& e) {
    const std::string what = e.what();
    if (m) qpb_multi_destroy(m);  // sets quit, joins the workers that did start, destroys every handle
    return qpb_internal_fail(QPB_ERR_NO_MEMORY, "qpb_multi_create: " + what);
  } catch (...) {
    if (m) qpb_multi_destroy(m);
    return qpb_internal_fail(QPB_ERR_NO_MEMORY, "qpb_multi_create: allocation failed");
  }
  *out = m;
  return QPB_SUCCESS;
}

int qpb_multi_destroy(qpb_multi_handle* m) {
  if (!m) return QPB_SUCCESS;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->quit = true;
  }
  m->cv_go.notify_all();
  int rc = QPB_SUCCESS;
  for (auto& w : m->workers) {
    if (w.thread.joinable()) w.thread.join();
    const int r = qpb_destroy(w.h);
    if (r != QPB_SUCCESS) rc = r;
  }
  delete m;
  return rc;
}

int qpb_multi_num_shards(const qpb_multi_handle* m) { return m ? (int)m->workers.size() : 0; }

int qpb_multi_set_joint_gains(qpb_multi_handle* m, const qpb_joint_gains* gains) {
  if (!m) return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_multi_set_joint_gains: null handle");
  for (auto& w : m->workers) {
    const int rc = qpb_set_joint_gains(w.h, gains);
    if (rc != QPB_SUCCESS) return rc;
  }
  return QPB_SUCCESS;
}

int qpb_multi_control_batch_host(qpb_multi_handle* m, int64_t n, const qpb_state_rec* h_states, qpb_out_rec* h_out) {
  if (!m || n < 0 || (n > 0 && (!h_states || !h_out)))
    return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_multi_control_batch_host: bad argument");
  if (n == 0) return QPB_SUCCESS;
  return run_sharded(m, n, h_states, nullptr, h_out);
}

int qpb_multi_tick_batch_host(qpb_multi_handle* m, int64_t n, const qpb_state_rec* h_states, const qpb_swing_rec* h_swing,
                              qpb_out_rec* h_out) {
  if (!m || n < 0 || (n > 0 && (!h_states || !h_swing || !h_out)))
    return qpb_internal_fail(QPB_ERR_INVALID_ARG, "qpb_multi_tick_batch_host: bad argument");
  if (n == 0) return QPB_SUCCESS;
  return run_sharded(m, n, h_states, h_swing, h_out);
}

int64_t qpb_multi_launch_count(const qpb_multi_handle* m) {
  int64_t total = 0;
  if (m)
    for (const auto& w : m->workers) total += qpb_launch_count(w.h);
  return total;
}

}  // extern "C"
