// epi_handle.h — the opaque handle behind include/epimcmc.h (host side only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/epimcmc.h"
#include "epi_device.cuh"

namespace epi {

// device scratch of one evaluation batch; grown on demand, never shrunk
struct Workspace {
  int64_t cap = 0, ld = 0;
  int P = 0, P1 = 0;
  int ns = 0;  // series per chain `series` was allocated for (0: none)
  double *theta = nullptr, *aos = nullptr;    // staged theta: SoA d x ld, and the AoS landing buffer
  double *hist1 = nullptr, *hist2 = nullptr;  // S histories, [chain][group][pitch]
  double *ckpt = nullptr;                     // pass-1 checkpoints, [chain][state][ckpt days]
  double *W = nullptr;                        // latency profiles, [chain][NK][EPI_MAX_TAPS]
  int2 *pmeta = nullptr;                      // (dmin, n) per profile
  int *offset = nullptr, *pad = nullptr, *status = nullptr;
  int *need_r = nullptr;  // chains whose ODE passes must be redone with the removed compartment integrated
  double *logl = nullptr, *logp = nullptr, *terms = nullptr, *series = nullptr;
};

struct SamplerState;

}  // namespace epi

struct epi_handle {
  int device = 0;
  epi_model_desc desc{};
  epi::DevModel M{};
  const epi::PriorTerm *PT = nullptr;
  int n_prior = 0;
  int n_terms = 0;
  double box_min[EPI_MAX_PARAMS]{}, box_max[EPI_MAX_PARAMS]{}, init[EPI_MAX_PARAMS]{};
  std::vector<double> sizes, size_vec;  // NB size arguments kept alive while the term table is built
  std::vector<void *> owned;            // device allocations freed at destroy
  epi::Workspace ws, ws_traj;
  epi::SamplerState *sampler = nullptr;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  // the sampler runs a small population as `ways` slices, each a free-running pipeline on its own stream
  // (epi_sampler_run); split_stream[k] serves slice k + 1, the handle's stream slice 0
  static constexpr int kMaxWays = 8;
  cudaStream_t split_stream[kMaxWays - 1] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxWays - 1] = {};
  int split_ways = 0;  // EPI_SPLIT_WAYS=1..8: force the number of slices (0: by population size)
  bool force_need_r = false;  // EPI_FORCE_R_FALLBACK=1: every chain takes the R-tracking instance (bit-identical results)
  bool pass1_chunk = true;  // EPI_PASS1_CHUNK=0: whole-period pass 1 in one round (bit-identical results)
  bool window_only = true;  // EPI_WINDOW_ONLY=0: pass 2 and the score sweep cover the whole period (bit-identical results)
  int ode_lanes = 0;  // EPI_ODE_LANES=1|2|4: force the lanes-per-chain mapping of the age ODE kernels (0: by batch size)
  int ode_minb = 0;   // EPI_ODE_MINB=4|5: force the register budget (blocks per SM) of the two-lane ODE kernel
  bool restart = true;  // pass 2 resumes from the pass-1 checkpoint (EPI_PASS2_RESTART=0: re-integrate everything)
  // epi_eval_submit/wait: two staging buffers for the raw host theta and the events that order
  // copy stream (stream2) and compute stream
  double *stage[2] = {nullptr, nullptr};
  int64_t stage_cap[2] = {0, 0};
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  bool slot_used[2] = {false, false}, slot_pending[2] = {false, false};
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // whole-batch evaluations share one scratch workspace: the stream the last one ran on and an event at its tail,
  // so that an evaluation on ANOTHER stream (epi_eval_device with a caller's stream) is ordered behind it
  cudaStream_t ws_stream = nullptr;
  cudaEvent_t ev_ws = nullptr;
  bool ws_busy = false;
  bool timing = false;
  int64_t launches = 0;
  std::string err;
};

int epi_fail(epi_handle *h, int rc, const char *fmt, ...);
void epi_sampler_free(epi_handle *h);
int epi_ensure_workspace(epi_handle *h, epi::Workspace &w, const epi::DevModel &M, int64_t n, int want_series);
int epi_launch_eval_range(epi_handle *h, epi::Workspace &w, const epi::DevModel &M, const double *d_theta, int64_t ld,
                          int64_t first, int64_t n, double *d_logl, double *d_logp, int *d_status, cudaStream_t stream);
int epi_launch_eval_ws(epi_handle *h, epi::Workspace &w, const epi::DevModel &M, const double *d_theta, int64_t ld,
                       int64_t n, int model_space, double *d_logl, double *d_logp, int *d_status, double *d_terms,
                       double *d_series, cudaStream_t stream, int ns_total = 0);
