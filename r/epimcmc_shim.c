/*
 * epimcmc_shim.c — R `.Call` shim over libepimcmc.so (include/epimcmc.h).
 *
 * SOURCE ONLY in this repository: R, R.h and Rinternals.h are not installed in the
 * build image, so this file is not built or run here; it is type-checked against
 * include/epimcmc.h with a declarations-only stand-in for R's C API (tests/r_stub,
 * tests/test_capi.py), and the same C entry points are driven through Python ctypes by
 * tests/.  On a machine with R:
 *
 *   R CMD SHLIB epimcmc_shim.c -I../include -L../epi_mcmc_b200/csrc -lepimcmc -o epimcmc_shim.so
 *
 * It replaces what a reference model file does at load time,
 *   system("R CMD SHLIB model-agen.cpp"); dyn.load("model-agen.so")      (model-age-groups2q.R:2-3)
 * and binds, one to one:
 *   C_epi_create      -> epi_create      handle from the data/settings globals
 *   C_epi_logpost     -> epi_eval        batched calclogl(transformParams(.)) + calclogp(.)
 *   C_epi_trajectories-> epi_trajectories / epi_trajectories_ext   calculateModel(params, period) for many draws
 *   C_epi_quantiles   -> epi_quantiles   quantileData (libMCMC.R:151-172)
 *   C_epi_marginalize -> epi_marginalize marginalizeData (libMCMC.R:174-188)
 *   C_epi_df_params   -> epi_df_params
 *   C_epi_run_mcmc    -> epi_sampler_*   drjacoby::run_mcmc (fitMCMC-drj.R:48-57) for a population of chains
 *                                        that stays on the GPU: burn-in, then recorded sampling draws
 * R is single-threaded and drjacoby calls back on the main thread, which is all the
 * handle's "one host thread at a time" rule needs.  Errors: the library never longjmps;
 * a nonzero return code is turned into Rf_error() AFTER every temporary is released.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "epimcmc.h"

static void handle_finalizer(SEXP ptr) {
  epi_handle *h = (epi_handle *)R_ExternalPtrAddr(ptr);
  if (h) {
    epi_destroy(h);
    R_ClearExternalPtr(ptr);
  }
}

static epi_handle *get_handle(SEXP ptr) {
  epi_handle *h = (epi_handle *)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("epimcmc: handle already destroyed");
  return h;
}

static double num(SEXP lst, const char *name, double dflt) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  if (names == R_NilValue) return dflt;  /* an unnamed list has nothing to look up */
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asReal(VECTOR_ELT(lst, i));
  return dflt;
}

static const double *vec(SEXP lst, const char *name, int *n) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  if (names == R_NilValue) { if (n) *n = 0; return NULL; }
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) {
      SEXP v = VECTOR_ELT(lst, i);
      if (!Rf_isReal(v)) Rf_error("epimcmc: `%s` must be a numeric (double) vector", name);
      if (n) *n = (int)XLENGTH(v);
      return REAL(v);
    }
  if (n) *n = 0;
  return NULL;
}

/* .Call(C_epi_create, flavour, period, substeps, globals, device): `globals` is a named list
 * holding the data-file and settings.R variables (same names with "." -> "_"). */
SEXP C_epi_create(SEXP flavour, SEXP period, SEXP substeps, SEXP g, SEXP device) {
  epi_model_desc desc = {Rf_asInteger(flavour), Rf_asInteger(period), Rf_asInteger(substeps), 0};
  epi_data D;
  memset(&D, 0, sizeof D);
  D.N = num(g, "N", 0); D.y_N = num(g, "y_N", 0); D.o_N = num(g, "o_N", 0);
  D.total_deaths_at_lockdown = num(g, "total_deaths_at_lockdown", 0);
  D.dmort_last = num(g, "dmort_last", 0);
  D.lockdown_offset = (int)num(g, "lockdown_offset", 0);
  D.lockdown_transition_period = (int)num(g, "lockdown_transition_period", 0);
  D.d3 = (int)num(g, "d3", 0); D.d4 = (int)num(g, "d4", 0); D.d5 = (int)num(g, "d5", 0);
  D.d6 = (int)num(g, "d6", 0); D.d7 = (int)num(g, "d7", 0); D.d8 = (int)num(g, "d8", 0);
  D.d_reliable_cases = (int)num(g, "d_reliable_cases", 0);
  D.d_reliable_hosp = (int)num(g, "d_reliable_hosp", 0);
  D.d_hosp_o1 = (int)num(g, "d_hosp_o1", 0); D.d_hosp_o2 = (int)num(g, "d_hosp_o2", 0);
  D.dhospi = vec(g, "dhospi", &D.n_dhospi);
  D.n_dhosp = (int)num(g, "n_dhosp", D.n_dhospi);
  D.dmorti = vec(g, "dmorti", &D.n_dmorti);
  D.y_dcasei = vec(g, "y_dcasei", &D.n_dcasei); D.o_dcasei = vec(g, "o_dcasei", NULL);
  int nm = 0;
  D.y_dmorti = vec(g, "y_dmorti", &nm);
  if (nm) D.n_dmorti = nm;
  D.o_dmorti = vec(g, "o_dmorti", NULL); D.o_wdmorti = vec(g, "o_wdmorti", NULL);
  D.y_ifr = vec(g, "y_ifr", &D.n_rate); D.o_ifr = vec(g, "o_ifr", NULL);
  D.y_hr = vec(g, "y_hr", NULL); D.o_hr = vec(g, "o_hr", NULL);
  D.g = vec(g, "g", NULL); D.gtrimp = vec(g, "gtrimp", NULL);
  D.hosp_nbinom_size = num(g, "hosp_nbinom_size", 0); D.mort_nbinom_size = num(g, "mort_nbinom_size", 0);
  D.case_nbinom_size1 = num(g, "case_nbinom_size1", 0); D.case_nbinom_sizey2 = num(g, "case_nbinom_sizey2", 0);
  D.case_nbinom_sizeo2 = num(g, "case_nbinom_sizeo2", 0); D.hosp_nbinom_size1 = num(g, "hosp_nbinom_size1", 0);
  D.hosp_nbinom_size2 = num(g, "hosp_nbinom_size2", 0); D.hosp_last7 = num(g, "hosp_last7", 0);
  /* age-2x only (R/models/model-age-groups2x.R:216-221,278-279,734,749) */
  D.d9 = (int)num(g, "d9", 0); D.d10 = (int)num(g, "d10", 0); D.d12 = (int)num(g, "d12", 0);
  D.duncertain = (int)num(g, "duncertain", 0);
  D.d_symp_cases = (int)num(g, "d_symp_cases", 0); D.d_all_cases = (int)num(g, "d_all_cases", 0);
  D.symp_cases_factor = num(g, "symp_cases_factor", 0);
  epi_handle *h = NULL;
  int rc = epi_create(&desc, &D, Rf_asInteger(device), &h);
  if (rc) Rf_error("epi_create failed (%d): %s", rc, epi_last_error(NULL));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* .Call(C_epi_logpost, handle, theta): theta is a d x n numeric matrix, one proposal per COLUMN
 * (R column-major == EPI_LAYOUT_AOS).  Returns list(logl, logp, status). */
SEXP C_epi_logpost(SEXP ptr, SEXP theta) {
  epi_handle *h = get_handle(ptr);
  const int d = epi_n_params(h);
  if (!Rf_isReal(theta) || XLENGTH(theta) % d != 0) Rf_error("epimcmc: theta must be a numeric %d x n matrix", d);
  const R_xlen_t n = XLENGTH(theta) / d;
  SEXP logl = PROTECT(Rf_allocVector(REALSXP, n)), logp = PROTECT(Rf_allocVector(REALSXP, n));
  SEXP status = PROTECT(Rf_allocVector(INTSXP, n));
  int rc = epi_eval(h, REAL(theta), (int64_t)n, EPI_LAYOUT_AOS, REAL(logl), REAL(logp), INTEGER(status));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, logl); SET_VECTOR_ELT(out, 1, logp); SET_VECTOR_ELT(out, 2, status);
  UNPROTECT(4);
  if (rc) Rf_error("epi_eval failed (%d): %s", rc, epi_last_error(h));
  return out;
}

/* .Call(C_epi_trajectories, handle, params, period, ext): params d x n MODEL-space; returns a
 * period x n_series x n array plus offset and padding.  ext = TRUE adds the state fields derived from the
 * other columns of the ode() matrix (y.E ... o.R, Re, Rt, y.Re, o.Re; epi_trajectories_ext). */
SEXP C_epi_trajectories(SEXP ptr, SEXP params, SEXP period, SEXP ext) {
  epi_handle *h = get_handle(ptr);
  const int want_ext = Rf_asLogical(ext) == TRUE;
  const int d = epi_n_params(h), ns = want_ext ? epi_n_series_ext(h) : epi_n_series(h), P = Rf_asInteger(period);
  if (!Rf_isReal(params) || XLENGTH(params) == 0 || XLENGTH(params) % d)
    Rf_error("epimcmc: params must be a numeric (double) matrix with %d rows", d);
  const R_xlen_t n = XLENGTH(params) / d;
  SEXP arr = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)P * ns * n));
  SEXP off = PROTECT(Rf_allocVector(INTSXP, n)), pad = PROTECT(Rf_allocVector(INTSXP, n));
  int rc = want_ext ? epi_trajectories_ext(h, REAL(params), (int64_t)n, EPI_LAYOUT_AOS, P, REAL(arr), INTEGER(off), INTEGER(pad))
                    : epi_trajectories(h, REAL(params), (int64_t)n, EPI_LAYOUT_AOS, P, REAL(arr), INTEGER(off), INTEGER(pad));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, arr); SET_VECTOR_ELT(out, 1, off); SET_VECTOR_ELT(out, 2, pad);
  UNPROTECT(4);
  if (rc) Rf_error("epi_trajectories failed (%d): %s", rc, epi_last_error(h));
  return out;
}

/* .Call(C_epi_quantiles, handle, params, period, startoffset, series_a, series_b, probs): quantileData of
 * R/lib/libMCMC.R:151-172 on the device.  params d x n MODEL-space (transformParams already applied);
 * series_a / series_b are 0-based indices into the ext series list (series_b = -1: a single series, else
 * their sum).  Returns a period x length(probs) matrix. */
SEXP C_epi_quantiles(SEXP ptr, SEXP params, SEXP period, SEXP startoffset, SEXP series_a, SEXP series_b, SEXP probs) {
  epi_handle *h = get_handle(ptr);
  const int d = epi_n_params(h), P = Rf_asInteger(period), nq = (int)XLENGTH(probs);
  if (!Rf_isReal(params) || XLENGTH(params) == 0 || XLENGTH(params) % d)
    Rf_error("epimcmc: params must be a numeric (double) matrix with %d rows", d);
  if (!Rf_isReal(probs) || nq < 1) Rf_error("epimcmc: probs must be a non-empty numeric (double) vector");
  const R_xlen_t n = XLENGTH(params) / d;
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, P, nq));
  /* the library returns period x n_probs row-major; R matrices are column-major */
  double *tmp = (double *)R_alloc((size_t)P * nq, sizeof(double));
  int rc = epi_quantiles(h, REAL(params), (int64_t)n, EPI_LAYOUT_AOS, P, Rf_asInteger(startoffset), Rf_asInteger(series_a),
                         Rf_asInteger(series_b), REAL(probs), nq, tmp);
  if (!rc)
    for (int j = 0; j < P; ++j)
      for (int q = 0; q < nq; ++q) REAL(out)[j + (R_xlen_t)q * P] = tmp[(size_t)j * nq + q];
  UNPROTECT(1);
  if (rc) Rf_error("epi_quantiles failed (%d): %s", rc, epi_last_error(h));
  return out;
}

/* .Call(C_epi_marginalize, handle, params, period, startoffset, series_a, series_b, kind, j1, j2): marginalizeData of
 * R/lib/libMCMC.R:174-188 for marginfun = max (kind 0) / sum (1) / min (2) / mean (3) of v[j1:j2]
 * (R/MCMC/predictMCMC.R:714-717 takes the maximum of Re over the next 60 days).  Returns one number per draw. */
SEXP C_epi_marginalize(SEXP ptr, SEXP params, SEXP period, SEXP startoffset, SEXP series_a, SEXP series_b, SEXP kind,
                       SEXP j1, SEXP j2) {
  epi_handle *h = get_handle(ptr);
  const int d = epi_n_params(h);
  if (!Rf_isReal(params) || XLENGTH(params) == 0 || XLENGTH(params) % d)
    Rf_error("epimcmc: params must be a numeric (double) matrix with %d rows", d);
  const R_xlen_t n = XLENGTH(params) / d;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  int rc = epi_marginalize(h, REAL(params), (int64_t)n, EPI_LAYOUT_AOS, Rf_asInteger(period), Rf_asInteger(startoffset),
                           Rf_asInteger(series_a), Rf_asInteger(series_b), Rf_asInteger(kind), Rf_asInteger(j1),
                           Rf_asInteger(j2), REAL(out));
  UNPROTECT(1);
  if (rc) Rf_error("epi_marginalize failed (%d): %s", rc, epi_last_error(h));
  return out;
}

SEXP C_epi_df_params(SEXP ptr) {
  epi_handle *h = get_handle(ptr);
  const int d = epi_n_params(h);
  SEXP mn = PROTECT(Rf_allocVector(REALSXP, d)), mx = PROTECT(Rf_allocVector(REALSXP, d));
  SEXP init = PROTECT(Rf_allocVector(REALSXP, d)), nm = PROTECT(Rf_allocVector(STRSXP, d));
  epi_df_params(h, REAL(mn), REAL(mx), REAL(init));
  for (int i = 0; i < d; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(epi_param_name(h, i)));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, nm); SET_VECTOR_ELT(out, 1, mn); SET_VECTOR_ELT(out, 2, mx); SET_VECTOR_ELT(out, 3, init);
  UNPROTECT(5);
  return out;
}

/* .Call(C_epi_run_mcmc, handle, theta0, chains, burnin, samples, thin, rungs, GTI_pow, seed, kind): what
 * fitMCMC-drj.R:48-57 asks of drjacoby::run_mcmc, for `chains` chains per rung on the device.
 *   theta0: NULL (start at init + jitter) or a d x (chains * rungs) numeric matrix of starting points;
 *   kind:   1 = DE-MC (population proposals), 0 = adaptive random-walk Metropolis, 2 = drjacoby's own update
 *           (one parameter at a time on the logit-reparameterised box, Robbins-Monro bandwidths; burnin / samples
 *           then count SWEEPS of d evaluations, as drjacoby's iterations do);
 *   rungs / GTI_pow: drjacoby's tempering ladder (1 = no tempering); draws are the COLD rung's.
 * Burn-in adapts the proposal shape from the population and refreshes the DE-MC partner snapshot every 10
 * iterations; both are frozen for the sampling phase.  Returns list(draws = d x chains x kept array,
 * logpost = chains x kept matrix, accept_rate, evals). */
SEXP C_epi_run_mcmc(SEXP ptr, SEXP theta0, SEXP chains, SEXP burnin, SEXP samples, SEXP thin, SEXP rungs, SEXP gti_pow,
                    SEXP seed, SEXP kind) {
  epi_handle *h = get_handle(ptr);
  const int d = epi_n_params(h), cpr = Rf_asInteger(chains), nr = Rf_asInteger(rungs) > 1 ? Rf_asInteger(rungs) : 1;
  const int n = cpr * nr, nburn = Rf_asInteger(burnin), nsamp = Rf_asInteger(samples);
  const int th = Rf_asInteger(thin) > 0 ? Rf_asInteger(thin) : 1, keep = nsamp / th > 0 ? nsamp / th : 1;
  if (cpr < 1 || nburn < 0 || nsamp < 1) Rf_error("epimcmc: chains >= 1, burnin >= 0, samples >= 1");
  if (theta0 != R_NilValue && (!Rf_isReal(theta0) || XLENGTH(theta0) != (R_xlen_t)d * n))
    Rf_error("epimcmc: theta0 must be NULL or a numeric %d x %d matrix", d, n);
  epi_sampler_cfg cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.kind = Rf_asInteger(kind) == 2 ? EPI_SAMPLER_DRJ : Rf_asInteger(kind) ? EPI_SAMPLER_DEMC : EPI_SAMPLER_RWM;
  cfg.n_chains = n;
  cfg.seed = (uint64_t)Rf_asReal(seed);
  cfg.scale = 1.0;
  cfg.de_eps = 1e-4;
  cfg.beta = 1.0;
  cfg.n_rungs = nr;
  cfg.gti_pow = Rf_asReal(gti_pow);
  int rc = epi_sampler_init(h, &cfg, theta0 == R_NilValue ? NULL : REAL(theta0));
  const int is_de = cfg.kind == EPI_SAMPLER_DEMC;
  if (!rc) rc = epi_sampler_adapt(h, 1, 0.0);
  for (int done = 0; !rc && done < nburn; done += 10) {
    rc = epi_sampler_run(h, nburn - done < 10 ? nburn - done : 10);
    if (!rc && is_de) rc = epi_sampler_set_population(h, NULL, 1, n, 0);
    R_CheckUserInterrupt();
  }
  if (!rc) rc = epi_sampler_adapt(h, 0, 0.0);
  if (!rc) rc = epi_sampler_record(h, th, keep);
  /* sampling phase: the partner snapshot taken at the end of burn-in stays FIXED (a fixed symmetric proposal
   * kernel; refreshing it from the moving chains would be an adaptive scheme without a valid archive) */
  int64_t it0 = 0, acc0 = 0, ev0 = 0;
  if (!rc) rc = epi_sampler_stats(h, &it0, &acc0, &ev0);
  for (int done = 0; !rc && done < nsamp; done += 10) {
    rc = epi_sampler_run(h, nsamp - done < 10 ? nsamp - done : 10);
    R_CheckUserInterrupt();
  }
  if (rc) Rf_error("epi_sampler failed (%d): %s", rc, epi_last_error(h));
  /* recorded draws of the cold rung = the LAST cpr chains, device layout kept x d x cpr */
  int32_t kept = 0;
  double *dr = (double *)R_alloc((size_t)keep * d * cpr, sizeof(double));
  double *lp = (double *)R_alloc((size_t)keep * cpr, sizeof(double));
  rc = epi_sampler_draws_subset(h, n - cpr, cpr, dr, lp, &kept);
  int64_t it = 0, acc = 0, ev = 0;
  if (!rc) rc = epi_sampler_stats(h, &it, &acc, &ev);
  if (rc) Rf_error("epi_sampler_draws failed (%d): %s", rc, epi_last_error(h));
  SEXP draws = PROTECT(Rf_alloc3DArray(REALSXP, d, cpr, kept)), logpost = PROTECT(Rf_allocMatrix(REALSXP, cpr, kept));
  for (int k = 0; k < kept; ++k)
    for (int p = 0; p < d; ++p)
      for (int c = 0; c < cpr; ++c)
        REAL(draws)[p + (R_xlen_t)d * (c + (R_xlen_t)cpr * k)] = dr[((size_t)k * d + p) * cpr + c];
  memcpy(REAL(logpost), lp, sizeof(double) * (size_t)kept * cpr);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, draws);
  SET_VECTOR_ELT(out, 1, logpost);
  /* acceptance of the SAMPLING phase (all rungs pooled; drjacoby reports it per phase too) */
  const double moves = (double)(it - it0) * n * (cfg.kind == EPI_SAMPLER_DRJ ? d : 1);  /* a sweep decides d moves per chain */
  SET_VECTOR_ELT(out, 2, Rf_ScalarReal(it > it0 ? (double)(acc - acc0) / moves : 0.0));
  SET_VECTOR_ELT(out, 3, Rf_ScalarReal((double)ev));
  UNPROTECT(3);
  return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_epi_create", (DL_FUNC)&C_epi_create, 5},
    {"C_epi_logpost", (DL_FUNC)&C_epi_logpost, 2},
    {"C_epi_trajectories", (DL_FUNC)&C_epi_trajectories, 4},
    {"C_epi_quantiles", (DL_FUNC)&C_epi_quantiles, 7},
    {"C_epi_marginalize", (DL_FUNC)&C_epi_marginalize, 9},
    {"C_epi_df_params", (DL_FUNC)&C_epi_df_params, 1},
    {"C_epi_run_mcmc", (DL_FUNC)&C_epi_run_mcmc, 10},
    {NULL, NULL, 0}};

void R_init_epimcmc_shim(DllInfo *dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
