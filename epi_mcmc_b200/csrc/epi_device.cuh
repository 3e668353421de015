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
constexpr int kMaxBp = 13;     // schedule breakpoints (model-agen.cpp:73-100: 10; model-ager.cpp:88-124: 13)
constexpr int kNaFlagTerms = 8;

// One slot of the day-major likelihood term table: the data-only part of one
// dnbinom() term that reads model day (dstart + rel).  (R recycling resolved on
// the host, see build_term_table in epi_capi.cu.)
struct Slot {
  double size;  // common size of the merged dnbinom() terms
  float x;      // sum of their observed counts (integers < 2^24: exact in a float)
  float n;      // how many terms were merged (0 => empty slot); sum of (x + size) = x + n * size
};
static_assert(sizeof(Slot) == 16, "one 16-byte load per slot");
__device__ __forceinline__ Slot ldg_slot(const Slot *p) {
  const double2 raw = __ldg(reinterpret_cast<const double2 *>(p));
  Slot s;
  s.size = raw.x;
  const long long b = __double_as_longlong(raw.y);
  s.x = __int_as_float((int)(b & 0xffffffffll));
  s.n = __int_as_float((int)(b >> 32));
  return s;
}

// Generalised beta(t) schedule of the two-age-group flavours, built on the host per model file
// (epi_capi.cu:build_schedule) — a new schedule is a new table, not new kernel code:
//   breakpoint k   t_k = ((data_offset + c) + A + B) + post,  A = theta[ia] (max(theta[ia], 0) when clamp), B = theta[ib]
//                  (ia / ib = -1: term absent; the operations run in exactly this order: the reference files build
//                  their breakpoints as running sums, e.g. t4 = (t2 + t3o) + t4o, model-age-groups2q.R:184-197)
//   beta triple k  beta_c = theta[i[c]] * (m[c] >= 0 ? theta[m[c]] : 1),  c = y, o, yo
//   segment k      linear towards k + 1 when bit k of `linear` is set, else constant (model-agen.cpp:78-99)
struct SchedBp { double c, post; short ia, ib, clamp, pad_; };
struct SchedBeta { short i[3], m[3]; };
struct Schedule {
  int n;
  unsigned linear;
  SchedBp bp[kMaxBp];
  SchedBeta beta[kMaxBp];
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
  // age-2x (model-age-groups2x.R + model-ager.cpp): waning rate, the fixed breakpoint days d9 / d12, the
  // symptomatic-testing factor window of the case series (:278-279,293-294)
  double eta, symp_factor;
  int d9, d12, d_symp, d_all;
  const Schedule *sched;  // device; age flavours: breakpoints, beta triples and segment kinds (read through the
                          // read-only cache: a by-value table would be copied to every thread's stack to be indexed)
  const double *y_ifr, *o_ifr, *y_hr, *o_hr, *g, *gtrimp;  // device, n_rate each
  const Slot *slots[kMaxSeries];                           // device, [n_obs][K]
  int multi_begin[kMaxSeries];  // first relative day with more than one slot (n_obs when K == 1)
  const double *box_min, *box_max;                         // device, d each (df_params min/max)
};

// One prior term: value = SUM: c0 + ca*theta[ia] + cb*theta[ib]; PROD: theta[ia]*theta[ib];
// RATIO: theta[ia]/theta[ib]; then dnorm(value, mean, sd, log) or dlnorm(value, mean, sd, log).
struct PriorTerm {
  int ia, ib, mode, is_ln;
  double c0, ca, cb, mean, sd;
  double log_sd;  // log(sd), filled on the host (data-only)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Regularised lower incomplete gamma P(a, x), x = q/scale, i.e. R's pgamma(q, shape, scale) of
// model-age-groups2q.R:31-32.  Same split as oracle_pgamma (power series below a+1, continued
// fraction above) but arranged for the GPU (ncu, profiles/r1_v1: the textbook loops are one long
// dependent chain through an FP64 division per term):
//   - series: four terms per trip, their four divisions independent of the running product;
//   - continued fraction: Wallis forward recurrence of numerator/denominator (two independent
//     FMA chains, no division inside the loop), convergence tested every four terms.
// Agreement with the oracle is a few ulp (tests: < 1e-13 relative on every profile weight).
// lga = lgamma(shape), hoisted by the caller because every CDF point of a profile shares it.
// METHOD: 0 = power series (x < a+1), 1 = continued fraction (x >= a+1).  The caller sorts the CDF
// points of a chain by method so that a warp round never executes both (ncu, profiles/r1_v3: with
// mixed rounds the two loops ran back to back for every round).
template <int METHOD>
__device__ __forceinline__ double dev_pgamma_m(double q, double shape, double scale, double lga) {
  const double x = q / scale, a = shape;
  if (!(x > 0.0)) return (x <= 0.0) ? 0.0 : x;  // NaN propagates
  const double pre = exp(-x + a * log(x) - lga);
  if constexpr (METHOD == 0) {
    // four terms per trip and ONE division for the four of them (ncu, profiles/r1_v7: the four
    // divisions x/(a+k) were 22 % of this kernel's instructions):
    //   d_k = del * x^k / ((ap+1)...(ap+k)) = del * x^k * [(ap+k+1)...(ap+4)] / [(ap+1)...(ap+4)]
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    double ap = a, del = 1.0 / a, sum = del;
    for (int n = 0; n < 25000; ++n) {
      const double a1 = ap + 1.0, a2 = ap + 2.0, a3 = ap + 3.0, a4 = ap + 4.0;
      const double p34 = a3 * a4, p234 = a2 * p34;
      const double s = del * (1.0 / (a1 * p234));  // the reciprocal does not depend on the running term
      const double d1 = s * x * p234, d2 = s * x2 * p34, d3 = s * x3 * a4, d4 = s * x4;
      sum += d1; sum += d2; sum += d3; sum += d4;
      del = d4;
      ap = a4;
      if (!(fabs(d4) >= fabs(sum) * 1e-17)) break;
    }
    return sum * pre;
  } else {
    // Q(a,x) = pre / (b0 + a1/(b1 + a2/(b2 + ...))), b_i = x + 2i + 1 - a, a_i = -i (i - a)
    double b = x + 1.0 - a;
    double Am1 = 1.0, A = b, Bm1 = 0.0, B = 1.0;  // A_{-1}, A_0, B_{-1}, B_0
    double hprev = 1.0 / b, h = hprev;
    double fi = 0.0;  // term index kept in FP64 (an int -> double conversion per term was 3.7 % of k_mid, r1_v11)
    for (int i = 1; i < 100000; i += 4) {  // convergence (one division) tested every four terms
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        fi += 1.0;
        const double an = -fi * (fi - a);
        b += 2.0;
        const double An = fma(b, A, an * Am1), Bn = fma(b, B, an * Bm1);
        Am1 = A; A = An; Bm1 = B; B = Bn;
      }
      // rescale check every 16 terms: one term grows |A| by at most |b| + |an| (< 1e10 even at the iteration
      // cap), so from 1e60 sixteen terms stay below 1e220
      if ((i & 15) == 13 && fabs(A) > 1e60) { A *= 1e-60; Am1 *= 1e-60; B *= 1e-60; Bm1 *= 1e-60; }
      h = B / A;
      if (!(fabs(h - hprev) > 1.2e-16 * fabs(h))) break;
      hprev = h;
    }
    return 1.0 - pre * h;
  }
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

// Table-driven natural logarithm for the negative-binomial terms (two logs per term, ~2,400 per
// evaluation: 29 % of the score kernel's instructions with the library log, ncu profiles/r1_v5).
// a = 2^e * m, m in [1,2); the top 7 mantissa bits pick c = 1 + (i + 1/2)/128 from a shared-memory
// table of (1/c, log c); r = m/c - 1 has |r| <= 2^-8, so log1p(r) needs six terms.  Absolute error
// ~2e-16 (the library's is 1 ulp); zero, denormal, negative, inf and NaN inputs take the library path.
constexpr int kLogTabN = 128;
__device__ __forceinline__ void fill_log_table(double2 *tab, int tid, int nthreads) {
  for (int i = tid; i < kLogTabN; i += nthreads) {
    const double c = 1.0 + ((double)i + 0.5) / (double)kLogTabN;
    tab[i] = make_double2(1.0 / c, log(c));
  }
}
__device__ __forceinline__ double tlog(double a, const double2 *__restrict__ tab) {
  const int hi = __double2hiint(a);
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(a);
  const int e = (hi >> 20) - 1023;
  const double2 t = tab[(hi >> 13) & (kLogTabN - 1)];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a));
  const double r = fma(m, t.x, -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  const double l = fma(r * r, p, r);
  return fma((double)e, 0.693147180559945309417232, t.y + l);
}

// the same for an argument known to be positive, finite and normal (the negative-binomial means are
// clamped to a positive floor and screened for NaN/inf once per term by the caller): no range branch
__device__ __forceinline__ double tlog_pos(double a, const double2 *__restrict__ tab) {
  const int hi = __double2hiint(a);
  const int e = (hi >> 20) - 1023;
  const double2 t = tab[(hi >> 13) & (kLogTabN - 1)];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a));
  const double r = fma(m, t.x, -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  const double l = fma(r * r, p, r);
  return fma((double)e, 0.693147180559945309417232, t.y + l);
}

// R pmax(floor, v): NA propagates (CUDA fmax would swallow the NaN)
__device__ __forceinline__ double dev_pmax(double lo, double v) { return (v < lo) ? lo : v; }

}  // namespace epi
