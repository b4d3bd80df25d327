// ubench_pcie.cu -- what the PCIe link of one GPU gives for the byte mix of the end-to-end path (512 B up, 256 B down per
// robot state): upload alone, download alone, both at once; default pinned memory against a write-combined input buffer.
// With argument N > 1 the same is done on N devices at once from N host threads (the host side is shared).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_pcie ubench_pcie.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)

struct Dev {
  int id;
  void *h_in = nullptr, *h_in_wc = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  cudaStream_t up, down;
};

static const size_t kIn = (size_t)65536 * 512, kOut = (size_t)65536 * 256;
static const int kReps = 40;

// mode: 1 = upload, 2 = download, 3 = both.  Returns seconds for kReps rounds on this device.
static double run(Dev& d, int mode, bool wc) {
  CK(cudaSetDevice(d.id));
  const void* src = wc ? d.h_in_wc : d.h_in;
  for (int w = 0; w < 3; w++) {
    if (mode & 1) CK(cudaMemcpyAsync(d.d_in, src, kIn, cudaMemcpyHostToDevice, d.up));
    if (mode & 2) CK(cudaMemcpyAsync(d.h_out, d.d_out, kOut, cudaMemcpyDeviceToHost, d.down));
  }
  CK(cudaStreamSynchronize(d.up));
  CK(cudaStreamSynchronize(d.down));
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < kReps; r++) {
    if (mode & 1) CK(cudaMemcpyAsync(d.d_in, src, kIn, cudaMemcpyHostToDevice, d.up));
    if (mode & 2) CK(cudaMemcpyAsync(d.h_out, d.d_out, kOut, cudaMemcpyDeviceToHost, d.down));
  }
  CK(cudaStreamSynchronize(d.up));
  CK(cudaStreamSynchronize(d.down));
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
  int have = 0;
  CK(cudaGetDeviceCount(&have));
  const int n = std::min(have, argc > 1 ? std::atoi(argv[1]) : 1);
  std::vector<Dev> devs(n);
  for (int i = 0; i < n; i++) {
    Dev& d = devs[i];
    d.id = i;
    CK(cudaSetDevice(i));
    CK(cudaHostAlloc(&d.h_in, kIn, cudaHostAllocDefault));
    CK(cudaHostAlloc(&d.h_in_wc, kIn, cudaHostAllocWriteCombined));
    CK(cudaHostAlloc(&d.h_out, kOut, cudaHostAllocDefault));
    std::memset(d.h_in, 1, kIn);
    std::memset(d.h_in_wc, 1, kIn);
    CK(cudaMalloc(&d.d_in, kIn));
    CK(cudaMalloc(&d.d_out, kOut));
    CK(cudaStreamCreateWithFlags(&d.up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d.down, cudaStreamNonBlocking));
  }
  const char* names[4] = { "", "upload only (512 B per state)", "download only (256 B per state)", "both at once" };
  for (int wc = 0; wc < 2; wc++)
    for (int mode = 1; mode <= 3; mode++) {
      if (wc && mode == 2) continue;
      std::vector<double> secs(n);
      std::vector<std::thread> th;
      for (int i = 0; i < n; i++) th.emplace_back([&, i] { secs[i] = run(devs[i], mode, wc != 0); });
      for (auto& t : th) t.join();
      const double worst = *std::max_element(secs.begin(), secs.end());
      const double bytes = (double)kReps * n * ((mode & 1 ? kIn : 0) + (mode & 2 ? kOut : 0));
      std::printf("%d GPU(s), %-31s %s: %7.1f GB/s in total = %.3e states/s\n", n, names[mode], wc ? "write-combined input" : "pinned input        ",
                  bytes / worst / 1e9, (double)kReps * n * 65536 / worst);
    }
  return 0;
}
