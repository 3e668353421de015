// Does a DFMA whose three sources are distinct 64-bit REGISTERS issue as fast as one with constant operands?
// K independent chains per thread; variant R: a[i] = fma(a[i], b[i], c[i]) with b, c in registers (loaded at run
// time); variant C: b, c compile-time constants (constant bank / immediates).  W warps per SM.
// cycles per DFMA per scheduler: 2.0 = the pipe's nominal rate (16 lanes per scheduler).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 dfma_operands.cu -o dfma_operands
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool REGS>
__global__ void k(double *out, const double *in, int iters, long long *cyc) {
  double a[K], b[K], c[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    a[i] = threadIdx.x * 1e-9 + i;
    b[i] = REGS ? in[i] : 1.0000001;
    c[i] = REGS ? in[K + i] : 1e-9;
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < K; ++i) a[i] = fma(a[i], b[i], c[i]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int K, bool REGS>
void run(int warps_per_sm, double *out, const double *in, long long *cyc) {
  const int iters = 1 << 14;
  k<K, REGS><<<148, warps_per_sm * 32>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double per_sched = warps_per_sm < 4 ? 1.0 : warps_per_sm / 4.0;
  printf("%s K=%2d warps/SM=%2d: %6.2f cycles per DFMA per scheduler\n", REGS ? "3 register sources" : "constant b, c      ", K,
         warps_per_sm, (double)h / iters / K / per_sched);
}

int main() {
  double *out, *in; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8); cudaMalloc(&in, 64 * sizeof(double));
  double hin[64];
  for (int i = 0; i < 64; ++i) hin[i] = i < 32 ? 1.0000001 + i * 1e-9 : 1e-9 * (i - 31);
  cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
  for (int w : {1, 4, 8, 16}) {
    run<8, false>(w, out, in, cyc); run<8, true>(w, out, in + 24, cyc);
    run<16, false>(w, out, in, cyc); run<16, true>(w, out, in + 16, cyc);
  }
  return 0;
}
