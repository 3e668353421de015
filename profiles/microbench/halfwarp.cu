// Does an FP64 instruction of a warp with only 16 active lanes occupy the SMSP's FP64 pipe (16 lanes / clk) for one
// cycle or for two?  8 independent DFMA chains per thread, W warps on ONE SM sub-partition (blockDim = 128 W puts W
// warps on each of the four), active lanes: all 32, the lower 16, the even 16.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 halfwarp.cu -o halfwarp
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double *out, int iters, long long *cyc, int mode, double b, double c) {
  const int lane = threadIdx.x & 31;
  const bool on = mode == 0 || (mode == 1 && lane < 16) || (mode == 2 && !(lane & 1)) || (mode == 3 && lane < 8);
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  __syncthreads();
  const long long t0 = clock64();
  if (on) {
    for (int i = 0; i < iters; ++i) {
      a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
      a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8 * 4096);
  const int iters = 20000;
  for (int W = 1; W <= 4; W *= 2)
    for (int mode = 0; mode < 4; ++mode) {
      k<<<1, 128 * W>>>(out, iters, cyc, mode, 1.0000001, 1e-9);
      k<<<1, 128 * W>>>(out, iters, cyc, mode, 1.0000001, 1e-9);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const char *names[] = {"32 lanes", "lower 16", "even 16", "lower 8"};
      printf("%d warps/SMSP, %s: %.3f cycles per warp-DFMA per SMSP\n", W, names[mode], (double)c / (8.0 * iters * W));
    }
  return 0;
}
