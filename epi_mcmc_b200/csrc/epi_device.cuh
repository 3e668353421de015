// epi_device.cuh — device-side model description and FP64 math helpers shared by
// the evaluation kernels (epi_kernels.cu) and the sampler (epi_sampler.cu).
// sm_100a only; no tensor-core work here: nothing on this path is a dense
// contraction, the bound is the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/epimcmc.h"

namespace epi {

constexpr int kTile = 32;      // chains per history tile == warp width
constexpr int kPadMax = 64;    // prefill slots kept in front of a staged S series
constexpr int kMaxSeries = 5;  // scored series: age y.casei,o.casei,hospi,y.deadi,o.deadi; simple hospi,deadi
constexpr int kMaxBp = 10;     // schedule breakpoints (model-agen.cpp:73-100)
constexpr int kNaFlagTerms = 8;

// One slot of the day-major likelihood term table: the data-only part of one
// dnbinom() term that reads model day (dstart + rel).  (R recycling resolved on
// the host, see build_term_table in epi_capi.cu.)
struct Slot {
  double x;   // observed count
  double xs;  // x + size  (0 => empty slot)
};

struct DevModel {
  int flavour, P, m, d;
  int P1;  // days integrated by pass 1 (60 for age-2q, P for simple/Ne)
  int lockdown_offset, ltp, d5, d8;
  int n_rate;
  int n_series;             // scored series
  int n_obs[kMaxSeries];    // observed length per scored series (window clamp)
  int K[kMaxSeries];        // slots per day per series
  int term_of[kMaxSeries];  // reference term index each series adds to
  int d_hosp_o1, d_hosp_o2;
  double N, Ny, No, a1, a2, gamma, a_inc;
  double ifr_y1, ifr_o1;  // y.ifr[1], o.ifr[1] (pass-1 death rates, model-age-groups2q.R:11-12)
  double floor_[kMaxSeries];
  double term_const[8];  // data-only constants per reference term (lgamma parts + size*log(size))
  double tdl, mort_size;  // simple: loglLD = dnbinom(total_deaths_at_lockdown, mu = max(.1, thr), size)
  double d3, d4, d5d, d6, d7, lo_ltp;
  const double *y_ifr, *o_ifr, *y_hr, *o_hr, *g, *gtrimp;  // device, n_rate each
  const Slot *slots[kMaxSeries];                           // device, [n_obs][K]
  const double *box_min, *box_max;                         // device, d each (df_params min/max)
};

// One prior term: value = SUM: c0 + ca*theta[ia] + cb*theta[ib]; PROD: theta[ia]*theta[ib];
// RATIO: theta[ia]/theta[ib]; then dnorm(value, mean, sd, log) or dlnorm(value, mean, sd, log).
struct PriorTerm {
  int ia, ib, mode, is_ln;
  double c0, ca, cb, mean, sd;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Regularised lower incomplete gamma P(a, x), x = q/scale: same algorithm and
// stopping rules as oracle_pgamma (series below a+1, Lentz continued fraction
// above), i.e. R's pgamma(q, shape, scale) of model-age-groups2q.R:31-32.
__device__ inline double dev_pgamma(double q, double shape, double scale) {
  const double x = q / scale, a = shape;
  if (!(x > 0.0)) return (x <= 0.0) ? 0.0 : x;  // NaN propagates
  const double lpre = -x + a * log(x) - lgamma(a);
  if (x < a + 1.0) {
    double ap = a, del = 1.0 / a, sum = del;
    for (int n = 0; n < 100000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(lpre);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; ++i) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - exp(lpre) * h;
}

__device__ __forceinline__ double dev_pnorm(double q, double mean, double sd) {
  return 0.5 * erfc(-((q - mean) / sd) * 0.70710678118654752440);
}

constexpr double kLnSqrt2Pi = 0.918938533204672741780329736406;

__device__ __forceinline__ double dev_dnorm_log(double x, double mean, double sd) {
  const double z = (x - mean) / sd;
  return -(kLnSqrt2Pi + 0.5 * z * z + log(sd));
}
__device__ __forceinline__ double dev_dlnorm_log(double x, double meanlog, double sdlog) {
  if (x != x) return x;
  if (x <= 0) return -INFINITY;
  const double y = (log(x) - meanlog) / sdlog;
  return -(kLnSqrt2Pi + 0.5 * y * y + log(x * sdlog));
}

// R pmax(floor, v): NA propagates (CUDA fmax would swallow the NaN)
__device__ __forceinline__ double dev_pmax(double lo, double v) { return (v < lo) ? lo : v; }

}  // namespace epi
