/*
 * epimcmc.h — C ABI of libepimcmc.so, the B200-native (sm_100a) batched
 * log-posterior evaluator + on-device sampler for the epi-mcmc model files.
 *
 * This is the drop-in boundary for ONE path of kdeforche/epi-mcmc: the
 * per-proposal log-posterior evaluation that the MCMC driver calls
 * (reference call sites, paths relative to the reference root):
 *
 *   R/MCMC/fitMCMC-drj.R:19-21   calcloglMCMC(params) = calclogl(transformParams(params))
 *   R/MCMC/fitMCMC-drj.R:48-57   run_mcmc(loglike = calcloglMCMC, logprior = calclogp, ...)
 *   R/models/model-age-groups2q.R:380-486  calclogp      (age flavour)
 *   R/models/model-age-groups2q.R:488-602  calclogl      (age flavour)
 *   R/models/model-age-groups2q.R:50-337   calculateModel (age flavour)
 *   R/models/model-simple-ode-drj-cmp.R:41-129,156-257   (simple flavour)
 *   R/models/model-agen.cpp:65-163  initmod()/derivs()   (deSolve compiled-model ABI)
 *
 * The reference's only native ABI is deSolve's per-RHS-call `derivs`
 * (model-agen.cpp:103-106), which is far too fine-grained for a GPU.  The
 * boundary therefore sits one level up, at the model-file API: an R `.Call`
 * shim (r/epimcmc_shim.c) binds exactly the entry points
 * declared here.
 *
 * Conventions: plain C, integer return codes (0 = ok), no exceptions cross the
 * ABI, caller owns every host buffer, the library owns every device buffer,
 * no global mutable state (contrast `static double parms[PARAM_N]`,
 * model-agen.cpp:4-5), one handle may be used by one host thread at a time.
 * There is NO CPU fallback: every compute entry point fails with
 * EPI_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef EPIMCMC_H
#define EPIMCMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPI_ABI_VERSION 2

/* ---- model flavours ------------------------------------------------------- */
/* simple 1-population SEIR, 8 parameters (R/models/model-simple-ode-drj-cmp.R) */
#define EPI_FLAVOUR_SIMPLE 0
/* simple + effective population size Nef, 9 parameters (README.md:78-95; the
 * model file itself is absent from the snapshot: restated from description)  */
#define EPI_FLAVOUR_NE 1
/* two-age-group SEEIR, 45 parameters (R/models/model-age-groups2q.R + model-agen.cpp) */
#define EPI_FLAVOUR_AGE2Q 2
/* two-age-group SEEIR, earlier generation: 36 parameters, 8-breakpoint schedule,
 * pass 1 over the whole period (R/models/model-age-groups2i.R + model-agef.cpp,
 * analyses/be/age/settings.R)                                                 */
#define EPI_FLAVOUR_AGE2I 3
/* two-age-group SEEIRS, later generation: 52 parameters, waning immunity (eta), effective-susceptible cap of the
 * younger group, 13-breakpoint schedule, the observation series are convolutions of the infection INCIDENCE that
 * derivs() reports as an output variable (not differences of N - filter(S)), a symptomatic-testing factor window
 * on the case series and two more "last 8 days" likelihood terms (R/models/model-age-groups2x.R + model-ager.cpp;
 * no analyses/ settings file of the snapshot points at this generation: its extra globals are epi_data members) */
#define EPI_FLAVOUR_AGE2X 4

#define EPI_MAX_PARAMS 56
/* widest latency profile supported: in-box maximum is 6*9+1 = 55 taps
 * (R/models/model-age-groups2q.R:22-23,656) */
#define EPI_MAX_TAPS 64
/* largest padding (= longest delay + 1) supported: age-2q/simple 64, age-2i 80 (its death
 * latency box reaches 50 + 3*9 = 77 days, R/models/model-age-groups2i.R:580-582) */
#define EPI_MAX_PADDING 80
/* InvalidDataOffset, R/models/model-age-groups2q.R:5 */
#define EPI_INVALID_DATA_OFFSET 10000

/* ---- return codes --------------------------------------------------------- */
#define EPI_OK 0
#define EPI_ERR_ARG 1        /* bad argument / inconsistent data description   */
#define EPI_ERR_NO_DEVICE 2  /* no usable CUDA device (there is no CPU path)   */
#define EPI_ERR_CUDA 3       /* a CUDA runtime call failed; see epi_last_error */
#define EPI_ERR_NOMEM 4
#define EPI_ERR_STATE 5      /* call sequence error (e.g. sampler not initialised) */

/* ---- per-evaluation status bits (int32 per proposal) ---------------------- */
/* pass 1 found no day with died > threshold: data_offset = 10000, pass 2 skipped
 * (R/models/model-age-groups2q.R:178-181,216,249-251,493-494)                 */
#define EPI_ST_INVALID_OFFSET 1
/* the reference would return NA: a hospital latency profile reaches into the
 * first `kend` filter outputs that stats::filter leaves NA
 * (R/models/model-age-groups2q.R:113-114 omits the hosp kernels from padding)  */
#define EPI_ST_NA 2
/* theta lies outside the df_params box (R/models/model-age-groups2q.R:637-663) */
#define EPI_ST_OUT_OF_BOX 8
/* a latency profile would need more than EPI_MAX_TAPS taps (or a padding beyond the
 * flavour's limit, see EPI_MAX_PADDING): result is NaN                        */
#define EPI_ST_UNSUPPORTED 16

/* ---- theta layouts -------------------------------------------------------- */
/* n proposals x d parameters, one proposal per row (C row-major) == an R
 * d x n matrix in column-major order (one proposal per column)               */
#define EPI_LAYOUT_AOS 0
/* d x n, one parameter per row: theta[p*n + i] (structure of arrays)          */
#define EPI_LAYOUT_SOA 1

/* ---- sampler kinds -------------------------------------------------------- */
#define EPI_SAMPLER_RWM 0 /* random-walk Metropolis, per-parameter Gaussian scale */
#define EPI_SAMPLER_DEMC 1 /* differential evolution MC (ter Braak 2006) on a population snapshot */
/* epi_sampler_cfg.reserved carries flags.  By default a DE-MC proposal moves a random subset of the coordinates
 * (DREAM-style crossover, CR drawn from {0.1, 0.25, 0.5, single coordinate}); this flag restores full-dimensional moves. */
#define EPI_SAMPLER_FLAG_FULLDIM 1
/* drjacoby's own update (the sampler R/MCMC/fitMCMC-drj.R:48-57 drives): one parameter at a time, a Gaussian step
 * of per-chain, per-parameter bandwidth on the logit-reparameterised box phi = log((theta - min) / (max - theta))
 * with the Jacobian term in the acceptance ratio, the bandwidths steered by Robbins-Monro towards a target
 * acceptance (0.44) while adaptation is on, n_rungs temperature rungs with beta_r = (r / (R-1))^GTI_pow and
 * Metropolis coupling of adjacent rungs after every sweep.  One epi_sampler_run iteration is one SWEEP over the
 * d parameters (d evaluations per chain), as one drjacoby iteration is; draws are recorded per sweep.           */
#define EPI_SAMPLER_DRJ 2
/* The n_rungs "rungs" are independent replicate populations, all at beta = 1 and never swapped: chains of one
 * replicate draw their DE-MC partners from that replicate only.  The spread of the replicate means is an honest
 * estimate of the Monte-Carlo error of the whole run (bench.py's ESS figure).                                   */
#define EPI_SAMPLER_FLAG_REPLICATES 2
/* DE-MC: by default a quarter of the proposals (the CR = 1 slot of the crossover mixture) moves a single coordinate,
 * chosen uniformly, with gamma = 2.38 / sqrt(2) times the scale mixture — one-parameter moves from the population's
 * own differences; this flag puts CR = 1 back into that slot.                                                     */
#define EPI_SAMPLER_FLAG_NO_SINGLE 4

typedef struct epi_handle epi_handle;

/* Model description: which model file, and the solver/horizon settings.
 * `period` is FitTotalPeriod (analyses/be/age2q/settings.R:31;
 * analyses/de/simple/settings.R:29).  `substeps` is the number of classical
 * RK4 sub-steps per day of the fixed-step integrator that replaces
 * deSolve::ode/LSODA (R/models/model-age-groups2q.R:163-165,211-213).         */
typedef struct epi_model_desc {
  int32_t flavour;
  int32_t period;
  int32_t substeps;
  int32_t reserved;
} epi_model_desc;

/* Everything the model files read from R globals (data file + settings.R),
 * snapshotted once at handle creation.  Unused members for a flavour may be
 * 0/NULL.  All series are double (R numeric); observed counts must be
 * non-negative integers stored as doubles.
 *
 * simple / Ne (R/models/model-simple-ode-drj-cmp.R:41-129,183-257):
 *   N, lockdown_offset, lockdown_transition_period, total_deaths_at_lockdown,
 *   dhospi[n_dhospi], dmorti[n_dmorti], dmort_last, hosp_nbinom_size,
 *   mort_nbinom_size
 * age-2q (R/models/model-age-groups2q.R, R/data/be/data.R:62-109,
 *         R/data/be/data-age-all.R:160-161,372-390, analyses/be/age2q/settings.R:29-38):
 *   y_N, o_N, lockdown_*, total_deaths_at_lockdown, d3..d8, d_reliable_cases,
 *   d_reliable_hosp, d_hosp_o1, d_hosp_o2, y_dcasei/o_dcasei[n_dcasei],
 *   dhospi[n_dhospi], n_dhosp (= length(dhosp)), y_dmorti/o_dmorti/o_wdmorti[n_dmorti],
 *   dmort_last, y_ifr/o_ifr/y_hr/o_hr/g/gtrimp[n_rate], case_nbinom_size1,
 *   case_nbinom_sizey2, case_nbinom_sizeo2, hosp_nbinom_size1, hosp_nbinom_size2,
 *   mort_nbinom_size, hosp_last7
 * age-2i (R/models/model-age-groups2i.R, analyses/be/age/settings.R:29-33): as age-2q minus
 *   g, gtrimp, o_wdmorti, d6, d8, d_reliable_hosp, d_hosp_o1/o2, hosp_nbinom_size1/2, hosp_last7;
 *   the settings' case_nbinom_size2 goes into BOTH case_nbinom_sizey2 and case_nbinom_sizeo2,
 *   hosp_nbinom_size and mort_nbinom_size are the scalar sizes of :496-523
 * age-2x (R/models/model-age-groups2x.R): as age-2q plus d9, d10, d12, duncertain, d_symp_cases, d_all_cases,
 *   symp_cases_factor (its settings file is not in the snapshot)
 */
typedef struct epi_data {
  double N, y_N, o_N;
  double total_deaths_at_lockdown;
  double dmort_last; /* dmort[length(dmort)], only used for the df_params box */
  int32_t lockdown_offset, lockdown_transition_period;
  int32_t d3, d4, d5, d6, d7, d8;
  int32_t d_reliable_cases, d_reliable_hosp, d_hosp_o1, d_hosp_o2;
  int32_t n_dhospi, n_dhosp, n_dmorti, n_dcasei, n_rate;
  int32_t reserved;
  const double *dhospi;
  const double *dmorti; /* simple: dmorti; age: unused */
  const double *y_dcasei, *o_dcasei;
  const double *y_dmorti, *o_dmorti, *o_wdmorti;
  const double *y_ifr, *o_ifr, *y_hr, *o_hr, *g, *gtrimp;
  double hosp_nbinom_size, mort_nbinom_size;
  double case_nbinom_size1, case_nbinom_sizey2, case_nbinom_sizeo2;
  double hosp_nbinom_size1, hosp_nbinom_size2, hosp_last7;
  /* age-2x only (R/models/model-age-groups2x.R:216-221,278-279,734,749): d9, d10, d12, duncertain,
   * d.symp.cases, d.all.cases, symp.cases.factor */
  int32_t d9, d10, d12, duncertain, d_symp_cases, d_all_cases;
  double symp_cases_factor;
} epi_data;

/* ---- lifecycle ------------------------------------------------------------ */
/* Replaces `system("R CMD SHLIB model-agen.cpp"); dyn.load("model-agen.so")`
 * + the data/settings globals (R/models/model-age-groups2q.R:2-3).
 * device = CUDA device ordinal.  On failure *out is NULL and the message is
 * available from epi_last_error(NULL).                                        */
int epi_create(const epi_model_desc *desc, const epi_data *data, int device, epi_handle **out);
void epi_destroy(epi_handle *h);
const char *epi_last_error(const epi_handle *h);
int epi_abi_version(void);

/* ---- model-file metadata (fit.paramnames / df_params / init) -------------- */
/* R/models/model-age-groups2q.R:604-663; model-simple-ode-drj-cmp.R:259-270   */
int epi_n_params(const epi_handle *h);
const char *epi_param_name(const epi_handle *h, int i);
int epi_df_params(const epi_handle *h, double *min, double *max, double *init);

/* ---- the hot path: batched calclogl(transformParams(theta)) + calclogp(theta)
 * Replaces n sequential calls of R/MCMC/fitMCMC-drj.R:19-21,51.
 * theta: HOST buffer, n x d in `layout`.  logl, logp: HOST, n doubles each
 * (either may be NULL).  status: HOST, n int32 (may be NULL).  Host<->device
 * copies happen inside the call.                                              */
int epi_eval(epi_handle *h, const double *theta, int64_t n, int layout,
             double *logl, double *logp, int32_t *status);

/* Asynchronous pair for back-to-back host batches.  epi_eval_submit queues the host->device
 * copy (on the handle's copy stream), the evaluation and the device->host copies of the results
 * and returns; epi_eval_wait blocks until that submission's results are in the host buffers.
 * Two submissions may be in flight (slot 0 and 1): the input copy of batch k+1 then overlaps the
 * kernels of batch k.  All host buffers must stay valid until the matching epi_eval_wait and
 * should be page-locked (pageable memory works but serialises the copy).  Same results as
 * epi_eval, bit for bit.                                                      */
int epi_eval_submit(epi_handle *h, int slot, const double *theta, int64_t n, int layout,
                    double *logl, double *logp, int32_t *status);
int epi_eval_wait(epi_handle *h, int slot);

/* Same, all buffers already resident on the handle's device.  theta_soa is
 * d x n (EPI_LAYOUT_SOA).  Launches on `stream` (a cudaStream_t passed as
 * void*, NULL = the handle's own stream) and does not synchronise the host.
 * Whole-batch evaluations of one handle share its scratch buffers: the library
 * orders each of them behind the previous one on the device (an event), on
 * whichever streams they run, so they never overlap; the caller orders the
 * call against whatever produces theta and consumes the results.             */
int epi_eval_device(epi_handle *h, const double *d_theta_soa, int64_t n,
                    double *d_logl, double *d_logp, int32_t *d_status, void *stream);

/* The chains [first, first + n) of a device-resident batch of n_total chains
 * (theta_soa d x ld, outputs indexed by chain like epi_eval_device's).  Ranges
 * that do not overlap use disjoint scratch rows and may be in flight on
 * different streams at the same time (the library does not order them; the
 * caller does, against the producer of theta and the consumer of the results).
 * Measured (profiles/r2_summary.md): slicing a latency-bound shard of <= 16 k
 * chains this way buys no throughput — each slice's ODE passes take as long as
 * the whole shard's.  `first` must be a multiple of 128; epi_reserve(h, n_total)
 * must have sized the workspace before the first range call (growing it would
 * free buffers other ranges are using).                                       */
int epi_eval_device_range(epi_handle *h, const double *d_theta_soa, int64_t ld, int64_t n_total,
                          int64_t first, int64_t n, double *d_logl, double *d_logp, int32_t *d_status,
                          void *stream);
int epi_reserve(epi_handle *h, int64_t n);

/* Per-evaluation detail for parity checks: offset[n] = state$offset
 * (data_offset or 10000), padding[n] = state$padding, terms[n*8] = the
 * likelihood terms in reference order (age: y.loglC, o.loglC, loglH,
 * loglHRatio, y.loglD, o.loglD — age-2i has no ratio term, its slot 3 is 0;
 * simple: loglLD, loglH, loglD), each the full dnbinom sum.  HOST buffers, any may be NULL.              */
int epi_eval_detail(epi_handle *h, const double *theta, int64_t n, int layout,
                    int32_t *offset, int32_t *padding, double *terms);

/* ---- calculateModel(params, period): trajectories ------------------------- */
/* Replaces R/models/model-age-groups2q.R:50-337 / libMCMC.R:151-172 callers.
 * theta is in MODEL space (already through transformParams).  out is HOST,
 * n x n_series x (period) doubles for days padding+1..padding+period, series
 * order: age: y.S,o.S,y.casei,o.casei,y.hospi,o.hospi,y.deadi,o.deadi (8);
 * simple: S,hospi,deadi (3).                                                  */
int epi_n_series(const epi_handle *h);
int epi_trajectories(epi_handle *h, const double *theta, int64_t n, int layout,
                     int32_t period, double *out, int32_t *offset, int32_t *padding);

/* Same, plus the state fields calculateModel derives from the other columns of the ode() matrix
 * (R/models/model-age-groups2q.R:216-227, model-simple-ode-drj-cmp.R:108-113) behind the basic
 * series: age (18): ... , y.E, y.I, y.R, o.E, o.I, o.R, Re, Rt, y.Re, o.Re; simple (8): ... , E, I,
 * R, Re, Rt.  Re..o.Re are derivs()' output variables (R/models/model-agen.cpp:159-163) at the
 * integer output times; where the reference's bounds guard returns before setting them they are
 * NaN.  epi_quantiles accepts these series indices too (e.g. quantileData(..., state$Re, ...),
 * R/MCMC/analyzeMCMC.R:265).                                                  */
int epi_n_series_ext(const epi_handle *h);
int epi_trajectories_ext(epi_handle *h, const double *theta, int64_t n, int layout,
                         int32_t period, double *out, int32_t *offset, int32_t *padding);

/* quantileData(sample, fun, startoffset, period, quantiles) of R/lib/libMCMC.R:151-172:
 * calculateModel(params, 3*period) for each of the n posterior draws (theta in
 * MODEL space), v = takeAndPad(fun(state), state$offset - startoffset, 3*period)
 * (libMCMC.R:26-36), then per day j = 1..period the R type-7 quantiles of v[j]
 * over the draws with na.rm = TRUE — all on the device (one sort per day).
 * fun(state) = series_a (+ series_b when >= 0), indices into the epi_trajectories
 * series order.  Draws with an invalid data_offset count as NA.  Up to 8,192 draws a day is sorted in shared memory,
 * beyond that in a global scratch buffer (period x 2^ceil(log2 n) doubles, at most 4 GiB).
 * out: HOST, period x n_probs, row-major (NaN where every draw is NA).          */
int epi_quantiles(epi_handle *h, const double *theta, int64_t n, int layout, int32_t period, int32_t startoffset,
                  int32_t series_a, int32_t series_b, const double *probs, int32_t n_probs, double *out);

/* marginalizeData(sample, fun, startoffset, period, marginfun) of R/lib/libMCMC.R:174-188: per posterior draw
 * v = takeAndPad(fun(state), state$offset - startoffset, 3*period) and ONE number marginfun(v), for the marginals
 * the reference computes (R/MCMC/predictMCMC.R:714-717 takes max(v[j1:j2]) of Re): kind 0 = max, 1 = sum, 2 = min,
 * 3 = mean of v[j1..j2] (1-based, inclusive, <= 3*period); an NA inside the range gives NA, as R's max()/sum() do.
 * theta, series_a/series_b as for epi_quantiles (no limit on n).  out: HOST, n doubles.                          */
#define EPI_MARGIN_MAX 0
#define EPI_MARGIN_SUM 1
#define EPI_MARGIN_MIN 2
#define EPI_MARGIN_MEAN 3
int epi_marginalize(epi_handle *h, const double *theta, int64_t n, int layout, int32_t period, int32_t startoffset,
                    int32_t series_a, int32_t series_b, int32_t kind, int32_t j1, int32_t j2, double *out);

/* ---- on-device sampler (chain state never leaves the GPU) ------------------ */
typedef struct epi_sampler_cfg {
  int32_t kind;        /* EPI_SAMPLER_* */
  int32_t n_chains;    /* chains on THIS device */
  int64_t chain_id0;   /* global id of local chain 0 (Philox stream = seed, chain id) */
  uint64_t seed;
  double scale;        /* RWM: proposal sd = scale * (max-min) per parameter;
                          DE-MC: gamma = scale * 2.38/sqrt(2d) */
  double de_eps;       /* DE-MC jitter sd, relative to (max-min) */
  double beta;         /* likelihood tempering exponent (1 = posterior) */
  int32_t n_rungs;     /* parallel tempering: temperature rungs (<= 1: none).  n_chains must be a
                          multiple; chain i sits on rung i / (n_chains / n_rungs); the LAST rung is the
                          cold one (beta = 1), as in drjacoby (fitMCMC-drj.R:53,62)                     */
  int32_t reserved;    /* flags: EPI_SAMPLER_FLAG_* (0 = defaults) */
  double gti_pow;      /* beta_r = ((r) / (n_rungs - 1)) ^ gti_pow, r = 0 .. n_rungs-1 (GTI_pow, :56)   */
} epi_sampler_cfg;

int epi_sampler_init(epi_handle *h, const epi_sampler_cfg *cfg, const double *theta0 /* HOST n_chains x d AOS, or NULL = init + jitter */);
/* run n_iter fused propose -> evaluate -> accept iterations */
int epi_sampler_run(epi_handle *h, int32_t n_iter);
/* Burn-in adaptation (switch it off before recording draws): while enabled the
 * proposal shape follows the population (RWM: Cholesky factor of the population
 * covariance; DE-MC: per-parameter sd of the jitter), re-measured at every
 * epi_sampler_run call.  target_accept in (0,1) additionally runs a per-chain
 * Robbins-Monro step-size multiplier towards that acceptance rate; <= 0 leaves
 * the classical step sizes (2.38/sqrt(d), 2.38/sqrt(2d)) alone — preferable on
 * this rugged target (ceil/floor kernel extents, integer data_offset).        */
int epi_sampler_adapt(epi_handle *h, int enable, double target_accept);
/* Replace the global step multiplier (epi_sampler_cfg.scale) between epi_sampler_run calls and redraw the
 * pending proposal with it.  The host front end uses it during burn-in to steer the POPULATION acceptance
 * rate (all ranks' chains together, so that results do not depend on the sharding) towards a target; it is
 * frozen for the sampling phase.                                                */
int epi_sampler_set_scale(epi_handle *h, double scale);
/* device pointers of the current chain states (d x ld SOA, ld >= n_chains) for
 * the NCCL all-gather of the DE-MC population; the pointers stay valid until
 * the next epi_sampler_init / epi_destroy                                      */
int epi_sampler_state_device(epi_handle *h, double **d_theta_soa, double **d_logl, double **d_logp, int64_t *ld);
/* install the gathered population snapshot the DE-MC proposals draw partners
 * from: caller-owned DEVICE memory laid out [rank][d][ld_rank] (exactly what an
 * all-gather of every rank's d x ld state block produces).  d_pop == NULL
 * refreshes the built-in single-rank snapshot from the current state.        */
int epi_sampler_set_population(epi_handle *h, const double *d_pop, int32_t n_ranks, int32_t chains_per_rank,
                               int64_t ld_rank);
/* copy the current state to HOST: theta n_chains x d AOS, logl, logp          */
int epi_sampler_get(epi_handle *h, double *theta, double *logl, double *logp);
/* parallel-tempering swap counters since init (adjacent-rung Metropolis swaps) */
int epi_sampler_pt_stats(epi_handle *h, int64_t *proposed, int64_t *accepted);
/* accumulated counters since init: iterations, accepted moves (summed over chains) */
int epi_sampler_stats(epi_handle *h, int64_t *n_iter, int64_t *n_accept, int64_t *n_evals);
/* record thinned draws on device: keep every `thin`-th state for up to
 * `capacity` draws per chain; epi_sampler_draws copies them out
 * (HOST, n_kept x n_chains x d)                                               */
int epi_sampler_record(epi_handle *h, int32_t thin, int32_t capacity);
int epi_sampler_draws(epi_handle *h, double *draws, double *logpost, int32_t *n_kept);
/* the recorded draws of chains [first, first + n_sub) only, in the device layout:
 * draws HOST n_kept x d x n_sub, logpost HOST n_kept x n_sub (either may be NULL) */
int epi_sampler_draws_subset(epi_handle *h, int32_t first, int32_t n_sub, double *draws, double *logpost, int32_t *n_kept);
/* the recorded draws where they lie: DEVICE pointers to [n_kept][d][ld] and [n_kept][ld] (valid until the next
 * epi_sampler_record / epi_sampler_init), for reductions over the whole population without a host copy         */
int epi_sampler_draws_device(epi_handle *h, double **d_draws, double **d_logpost, int32_t *n_kept, int64_t *ld);
/* EPI_SAMPLER_DRJ: the current proposal bandwidths (HOST, n_chains x d, phi scale) */
int epi_sampler_bandwidths(epi_handle *h, double *bw);

/* ---- measurement helpers --------------------------------------------------- */
/* DFMA-chain microbenchmark on the handle's device: measured FP64 TFLOP/s, the
 * roofline denominator for the evaluation kernels (MEASURED_PEAKS.json has no
 * FP64 entry).                                                                */
int epi_fp64_peak(epi_handle *h, double *tflops);
/* device time (ms) of each kernel of the most recent epi_eval/epi_eval_device
 * when timing is enabled: [0]=pass-1 ODE, [1]=profiles+threshold, [2]=pass-2
 * ODE, [3]=score.                                                             */
int epi_set_timing(epi_handle *h, int enable);
int epi_last_kernel_ms(epi_handle *h, float ms[4]);
/* The fused accept/propose kernel (K2) alone: `iters` launches on a synthetic population of n_chains chains
 * (uniform states in the box, log densities set so that a fixed quarter of the chains accepts every proposal, own-population
 * snapshot for DE-MC), timed with CUDA events on the handle's stream.  No evaluation kernels run and no K1
 * workspace is allocated, so n_chains may be far beyond what a sampler would hold (the HBM roofline of K2 needs
 * a population outside L2: 2^22 chains).  us_per_iter: mean device time of one launch; bytes_per_iter: the
 * ALGORITHMIC bytes of one launch (each chain: theta read and proposal written once, the partner values of the
 * moving coordinates, the accepted rows copied, 36 B of log densities and status) at the acceptance rate the
 * launches actually had (accept_rate) and the expected number of moving coordinates of the crossover scheme.   */
int epi_k2_bench(epi_handle *h, int32_t kind, int64_t n_chains, int32_t iters, double *us_per_iter, double *bytes_per_iter,
                 double *accept_rate);
/* number of kernel launches issued by this handle since creation             */
int64_t epi_launch_count(const epi_handle *h);
/* diagnostic: how many chains of the most recent evaluation batch (epi_eval* / sampler) had their ODE passes
 * recomputed by the R-tracking kernel instance (0 in any realistic regime); < 0 on error */
int64_t epi_fallback_count(epi_handle *h);
/* test hook: raw Philox4x32-10 words generated on the device,
 * out[4*j..4*j+3] = philox(counter = (j, 0, iter, use), key = seed)           */
int epi_philox_probe(epi_handle *h, uint64_t seed, uint32_t iter, uint32_t use, int32_t n, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* EPIMCMC_H */
