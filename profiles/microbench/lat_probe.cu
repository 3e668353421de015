// Dependent-issue latencies on sm_100a that decide the shape of the ODE kernels: DFMA / DMUL / DADD chains,
// SHFL.IDX of a double (two 32-bit shuffles), VOTE (ballot), and the mixed chain of one four-lane RK4 stage.
// One warp per SM unless stated.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 lat_probe.cu -o lat_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double *out, int iters, long long *cyc, double b, double c) {
  double a = threadIdx.x * 1e-9 + 1.0, a2 = a + 0.5;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) a = fma(a, b, c);
    if (OP == 1) a = __dmul_rn(a, b);
    if (OP == 2) a = __dadd_rn(a, c);
    if (OP == 3) a = __shfl_sync(0xffffffffu, a, lane ^ 1);
    if (OP == 4) { unsigned m = __ballot_sync(0xffffffffu, a > 0.5); a = __hiloint2double(__double2hiint(a), (int)m); }
    if (OP == 5) {  // one stage of the four-lane kernel: 2 shuffles -> mul -> fma -> mul -> select -> fma -> fma
      const double pv = __shfl_sync(0xffffffffu, a, lane ^ 1), iy = __shfl_sync(0xffffffffu, a, (lane & ~3) | 1);
      const double inf = __dmul_rn(fma(b, pv, __dmul_rn(c, iy)), a2);
      const double w = (lane & 1) ? a2 : inf;
      const double dv = fma(c, a, w), du = fma(b, pv, -w);
      a = fma(1e-9, dv, a); a2 = fma(1e-9, du, a2);
    }
    if (OP == 6) {  // the thread-per-chain stage's critical path: mul -> mul -> fma -> fma
      const double inf = __dmul_rn(__dmul_rn(b, a), a2);
      const double dv = fma(c, a, inf);
      a = fma(1e-9, dv, a); a2 = fma(1e-9, -inf, a2);
    }
    if (OP == 7) { int x = __double2hiint(a); x = __shfl_sync(0xffffffffu, x, lane ^ 1); a = __hiloint2double(x, __double2loint(a)); }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + a2;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, int warps, double *out, long long *cyc) {
  const int iters = 1 << 14;
  k<OP><<<148, warps * 32>>>(out, iters, cyc, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("%-46s warps/SM=%2d: %7.2f cycles per iteration\n", name, warps, (double)h / iters);
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16}) {
    run<0>("DFMA chain", w, out, cyc);
    run<1>("DMUL chain", w, out, cyc);
    run<2>("DADD chain", w, out, cyc);
    run<3>("SHFL.IDX of a double (2 x 32 bit) chain", w, out, cyc);
    run<7>("SHFL.IDX 32 bit chain", w, out, cyc);
    run<4>("VOTE.ANY (ballot) chain", w, out, cyc);
    run<5>("four-lane RK4 stage (2 shfl, 6 FP64, select)", w, out, cyc);
    run<6>("thread-per-chain stage critical path (4 FP64)", w, out, cyc);
  }
  return 0;
}
