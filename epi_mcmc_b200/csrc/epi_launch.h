// epi_launch.h — host-visible launch interface between epi_capi.cu and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "epi_device.cuh"

namespace epi {

struct EvalArgs {
  const DevModel *M;
  const PriorTerm *PT;  // device
  int n_prior;
  const double *theta;  // device SoA, d x ld
  int64_t ld, n;
  int model_space;      // theta already through transformParams (calculateModel callers)
  double *hist1, *hist2, *W;
  double *ckpt;  // pass-1 day-boundary states for the pass-2 restart (null: re-integrate from day 0)
  int2 *pmeta;
  int *offset, *pad, *status;
  double *logl, *logp, *terms;  // device, any may be null
  double *series_out;           // non-null: write calculateModel series instead of scoring
  int ns_total = 0;             // series per chain in series_out: 8/3 basic, 18/8 with the ext series
  int *need_r = nullptr;        // per chain: the hot ODE instance could not decide the R part of the bounds guard
  int force_need_r = 0;         // EPI_FORCE_R_FALLBACK=1 (tests): flag every chain
  int64_t *n_launches = nullptr;  // incremented by the number of kernels enqueued
  int pass1_chunk = 1;          // whole-period pass 1 in two rounds (EPI_PASS1_CHUNK=0: one round over all days)
  int ode_lanes = 0;            // lanes per chain of the age-flavour ODE kernels: 0 = by batch size, 1 | 2 | 4 forced (EPI_ODE_LANES)
  int ode_minb = 0;             // two-lane ODE kernel: 0 = by batch size, 4 | 5 = blocks per SM its instance is built for (EPI_ODE_MINB)
  int window_only = 1;          // scoring path: integrate / sweep only up to the data windows (EPI_WINDOW_ONLY=0: whole period)
  cudaStream_t stream;
  cudaEvent_t *ev;  // null or 5 events bracketing the 4 launches
  cudaEvent_t ev_after_mid = nullptr;  // optional: recorded after k_mid (stagger point of the two-stream mode)
};

cudaError_t launch_eval(const EvalArgs &a);
cudaError_t launch_aos_to_soa(const double *aos, double *soa, int64_t n, int d, int64_t ld, cudaStream_t s);
cudaError_t launch_quantiles(const double *series, const int *offset, const int *pad, int64_t n, int NS, int simperiod,
                             int period, int startoffset, int sa, int sb, double prefill_a, double prefill_b,
                             const double *d_probs, int n_probs, double *d_out, double *d_scratch, cudaStream_t s);
cudaError_t launch_marginal(const double *series, const int *offset, const int *pad, int64_t n, int NS, int simperiod,
                            int startoffset, int sa, int sb, double prefill_a, double prefill_b, int kind, int j1, int j2,
                            double *d_out, cudaStream_t s);
cudaError_t launch_dfma(double *out, int blocks, int threads, int iters, cudaStream_t s);

}  // namespace epi
