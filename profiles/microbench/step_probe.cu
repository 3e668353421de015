// Where do the cycles of one two-lane RK4 step go?  The arithmetic of k_ode_age2's regular sub-step (rhs2w / rk4_try2w
// of epi_kernels.cu, HOLD instance) in isolation, one warp per scheduler and more, with the pieces switched off one by
// one: SHFL = the I_y shuffle, GUARD = the integer-pipe bounds flag, VOTE = the per-step ballot + any (guard verdict)
// and the all(hold) vote.  Prints cycles per RK4 step.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 step_probe.cu -o step_probe
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned kFull = 0xffffffffu;
struct L2 { double a1, gamma; int shift; unsigned nb; };
struct S2 { double bA, bB; bool hold, fast_ok; };
__device__ __forceinline__ unsigned below_one(double S) {
  const unsigned v = (unsigned)__double2hiint(S) - 0x3ff00000u;
  return v | (0x40000000u - v);
}
__device__ __forceinline__ unsigned oob4(double S, double a, double b, double c, unsigned nhi_m1) {
  const unsigned mx = max(max((unsigned)__double2hiint(a), (unsigned)__double2hiint(b)), (unsigned)__double2hiint(c));
  return (nhi_m1 - mx) | below_one(S);
}
template <bool SHFL, bool GUARD>
__device__ __forceinline__ void rhs(const L2 &L, const S2 &sg, const double *y, double *dy, unsigned &flag) {
  const double S = y[0], E1 = y[1], E2 = y[2], I = y[3];
  const double iy = SHFL ? __shfl_sync(kFull, I, L.shift) : I;
  if (GUARD) flag |= oob4(S, E1, E2, I, L.nb);
  const double inf = __dmul_rn(fma(sg.bA, I, __dmul_rn(sg.bB, iy)), S);
  dy[0] = -inf;
  dy[1] = fma(-L.a1, E1, inf);
  dy[2] = fma(L.a1, E1, -E2);
  dy[3] = fma(-L.gamma, I, E2);
}
template <bool SHFL, bool GUARD>
__device__ __forceinline__ unsigned step(const L2 &L, const S2 &sg, const double *y, double *yn, double hh, double hh2, double h6) {
  double k[4], acc[4], yt[4];
  unsigned flag = 0;
  rhs<SHFL, GUARD>(L, sg, y, k, flag);
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i] = k[i]; yt[i] = fma(hh2, k[i], y[i]); }
  rhs<SHFL, GUARD>(L, sg, yt, k, flag);
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh2, k[i], y[i]); }
  rhs<SHFL, GUARD>(L, sg, yt, k, flag);
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh, k[i], y[i]); }
  rhs<SHFL, GUARD>(L, sg, yt, k, flag);
#pragma unroll
  for (int i = 0; i < 4; ++i) yn[i] = fma(h6, __dadd_rn(acc[i], k[i]), y[i]);
  return flag;
}
template <bool SHFL, bool GUARD, bool VOTE>
__global__ void probe(double *out, const double *in, int iters, long long *cyc) {
  const int lane = threadIdx.x & 31, g = lane & 1;
  L2 L; L.a1 = in[0]; L.gamma = in[1]; L.shift = lane & ~1; L.nb = (unsigned)__double2hiint(in[2]) - 1u;
  S2 sg; sg.bA = in[3] * (g ? 0.9 : 1.0); sg.bB = g ? in[4] : 0.0; sg.hold = in[5] > 0; sg.fast_ok = in[5] > 0;
  double y[4] = {g ? in[2] : in[2] - 1.0, g ? 0.0 : 1.0, 0.0, 0.0}, z[4];
  const double h = in[6], hh2 = 0.5 * h, h6 = h / 6.0;
  int bad = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (VOTE) {
      if (!__all_sync(kFull, sg.hold)) ++bad;
      unsigned f = step<SHFL, GUARD>(L, sg, y, z, h, hh2, h6);
      bool ok = ((__ballot_sync(kFull, (int)f < 0) >> L.shift) & 3u) == 0u && sg.fast_ok;
      if (__any_sync(kFull, !ok)) ++bad;
      if (!__all_sync(kFull, sg.hold)) ++bad;
      f = step<SHFL, GUARD>(L, sg, z, y, h, hh2, h6);
      ok = ((__ballot_sync(kFull, (int)f < 0) >> L.shift) & 3u) == 0u && sg.fast_ok;
      if (__any_sync(kFull, !ok)) ++bad;
    } else {
      unsigned f = step<SHFL, GUARD>(L, sg, y, z, h, hh2, h6);
      f |= step<SHFL, GUARD>(L, sg, z, y, h, hh2, h6);
      bad += (int)f < 0;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = y[0] + y[1] + y[2] + y[3] + bad;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <bool SHFL, bool GUARD, bool VOTE>
void run(const char *name, int warps_per_sm, double *out, const double *in, long long *cyc) {
  const int iters = 2000;
  probe<SHFL, GUARD, VOTE><<<148, warps_per_sm * 32>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM=%2d: %7.1f cycles per RK4 step (per warp)\n", name, warps_per_sm, (double)h / iters / 2);
}
int main() {
  double *out, *in; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&in, 64); cudaMalloc(&cyc, 8);
  const double hin[8] = {1.0 / 3.0, 1.0 / 1.4, 8.55e6, 0.9 / 8.55e6, 0.3 / 8.55e6, 1.0, 0.25, 0};
  cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
  for (int w : {4, 8, 16}) {
    run<true, true, true>("shfl + guard + votes", w, out, in, cyc);
    run<true, true, false>("shfl + guard", w, out, in, cyc);
    run<true, false, false>("shfl", w, out, in, cyc);
    run<false, true, false>("guard, no shfl", w, out, in, cyc);
    run<false, false, false>("arithmetic only", w, out, in, cyc);
  }
  return 0;
}
