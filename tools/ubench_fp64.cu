// Developer micro-benchmark: dependent-issue latency of DFMA / DMUL / MUFU.RCP64H / LDS.64, barrier cost and DFMA
// throughput per SM (independent accumulators, 8 and 16 warps) on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int iters) {
  __shared__ double sm[256];
  sm[threadIdx.x] = 1.0 + threadIdx.x * 1e-9;
  __syncthreads();
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64();
  double m = a;
  for (int i = 0; i < iters; i++) { m = m * b; m = m * b; m = m * b; m = m * b; }
  long long t2 = clock64();
  double r = m;
  for (int i = 0; i < iters; i++) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r)); r = y; }
  long long t3 = clock64();
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i++) { idx = (int)sm[idx & 255] + (idx & 1); }
  long long t4 = clock64();
  for (int i = 0; i < iters; i++) __syncthreads();
  long long t5 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
  out[threadIdx.x + 1] = a + m + r + idx;
}
// 16 independent accumulators per thread, 2 multiplier operands shared (like the sweep's rank-1 tile update)
__global__ void thr(double* out, long long* cyc, int iters) {
  double acc[16];
  for (int i = 0; i < 16; i++) acc[i] = out[i] + threadIdx.x;
  double x0 = out[20] + 1e-9, x1 = out[21] + 2e-9, y0 = out[22] + 1.0, y1 = out[23] + 1.0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      acc[i] = fma(-x0, y0, acc[i]);
      acc[i + 1] = fma(-x0, y1, acc[i + 1]);
      acc[i + 2] = fma(-x1, y0, acc[i + 2]);
      acc[i + 3] = fma(-x1, y1, acc[i + 3]);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 16; i++) s += acc[i];
  out[32 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 65536); cudaMalloc(&c, 64); cudaMemset(d, 0, 65536);
  long long h[5];
  for (int threads : {32, 256}) {
    lat<<<1, threads>>>(d, c, 1000); cudaDeviceSynchronize();
    cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost);
    printf("threads=%d: DFMA %.1f cyc, DMUL %.1f, RCP64H %.1f, LDS+cvt %.1f, BAR %.1f\n", threads, h[0] / 4000.0, h[1] / 4000.0, h[2] / 1000.0, h[3] / 1000.0, h[4] / 1000.0);
  }
  for (int threads : {128, 256, 512, 1024}) {
    thr<<<1, threads>>>(d, c, 2000); cudaDeviceSynchronize();
    cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
    const double fmas = 2000.0 * 16 * threads;
    printf("throughput threads=%d: %.1f DFMA lanes per clk per SM (%.2f cyc per warp-DFMA per SMSP)\n", threads, fmas / h[0], h[0] / (2000.0 * 16 * (threads / 128.0)));
  }
  return 0;
}
