// ubench_launch.cu -- the floor under a one-robot call: what a launch + wait costs on this box before any solving.
//   (a) empty kernel, cudaStreamSynchronize
//   (b) kernel reads a 512-B record from pinned mapped host memory and writes 256 B back, cudaStreamSynchronize
//   (c) the same, the host waits on a flag the kernel writes into mapped memory after a system fence (no driver call)
// Mean microseconds per call over 5000 calls.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_launch ubench_launch.cu
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>

__global__ void empty_kernel() {}

__global__ void echo_kernel(const double* in, double* out, volatile unsigned* flag, unsigned seq) {
  const int t = threadIdx.x;  // 32 threads
  double2 v = reinterpret_cast<const double2*>(in)[t];
  if (t < 16) reinterpret_cast<double2*>(out)[t] = make_double2(v.x + 1.0, v.y + 1.0);
  if (flag) {
    __syncwarp();
    __threadfence_system();
    if (t == 0) *flag = seq;
  }
}

template <class F>
static double mean_us(F&& f, int iters) {
  for (int i = 0; i < 500; i++) f();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++) f();
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
}

int main() {
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  double* h = nullptr;
  cudaHostAlloc(reinterpret_cast<void**>(&h), 4096, cudaHostAllocMapped);
  double* d = nullptr;
  cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0);
  for (int i = 0; i < 512; i++) h[i] = i;
  volatile unsigned* hflag = reinterpret_cast<volatile unsigned*>(h + 256);
  unsigned* dflag = reinterpret_cast<unsigned*>(d + 256);
  unsigned seq = 0;
  const int iters = 5000;
  const double a = mean_us([&] { empty_kernel<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); }, iters);
  const double b = mean_us([&] { echo_kernel<<<1, 32, 0, st>>>(d, d + 64, nullptr, 0u); cudaStreamSynchronize(st); }, iters);
  const double c = mean_us([&] {
    seq++;
    echo_kernel<<<1, 32, 0, st>>>(d, d + 64, dflag, seq);
    while (*hflag != seq) {}
  }, iters);
  cudaStreamSynchronize(st);
  std::printf("{\"empty_launch_sync_us\": %.2f, \"mapped_read_write_sync_us\": %.2f, \"mapped_read_write_flag_us\": %.2f, \"check\": %.1f}\n",
              a, b, c, h[64]);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
