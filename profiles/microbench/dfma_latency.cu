// DFMA issue/latency microbenchmark (sm_100a): K independent FMA chains per thread, W warps per SM.
// cycles per DFMA per warp-scheduler tell the dependent-issue latency (K = 1, 1 warp/scheduler) and how many
// independent DFMAs must be in flight to fill the FP64 pipe.  Build: nvcc -arch=sm_100a -O3 dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void k(double *out, int iters, long long *cyc) {
  double a[K];
#pragma unroll
  for (int i = 0; i < K; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < K; ++i) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int K>
void run(int warps_per_sm, double *out, long long *cyc) {
  const int iters = 1 << 14;
  int nsm = 148;
  k<K><<<nsm, warps_per_sm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double per_sched = warps_per_sm / 4.0;
  // DFMA warp-instructions issued per scheduler = iters * K * warps/scheduler
  printf("K=%d warps/SM=%2d: %.2f cycles per DFMA per warp, %.3f DFMA/clk/scheduler (peak 0.5)\n", K, warps_per_sm,
         (double)h / iters / K, iters * (double)K * per_sched / (double)h);
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 12, 16, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
  return 0;
}
