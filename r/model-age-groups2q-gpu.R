## model-age-groups2q-gpu.R — drop-in for R/models/model-age-groups2q.R.
##
## SOURCE ONLY here (no R in the build image).  Use it exactly like the reference model
## file: set `fitmodel` in settings.R to this file; fitMCMC-drj.R then sources data ->
## this file -> libfit.R unchanged.  Everything the reference file defines keeps its name
## and meaning (fit.paramnames, keyparamnames, init, df_params, transformParams,
## invTransformParams, calcNominalState, calclogp, calclogl, calculateModel); only the
## bodies of calclogl / calclogp / calculateModel call the GPU through the .Call shim, and
## the batched entry points the GPU is for are added next to them.

dyn.load("epimcmc_shim.so")            # replaces: system("R CMD SHLIB model-agen.cpp"); dyn.load("model-agen.so")

InvalidDataOffset <- 10000
G <- 4.7; Tinc1 <- 3; Tinc2 <- 1
gamma <- 1/((G - (Tinc1 + Tinc2)) * 2)

if (!exists("epi_substeps")) epi_substeps <- 4        # RK4 sub-steps per day of the fixed-step integrator
if (!exists("epi_device")) epi_device <- 0

epi_globals <- list(
    y_N = y.N, o_N = o.N, total_deaths_at_lockdown = total_deaths_at_lockdown,
    dmort_last = dmort[length(dmort)],
    lockdown_offset = lockdown_offset, lockdown_transition_period = lockdown_transition_period,
    d3 = d3, d4 = d4, d5 = d5, d6 = d6, d7 = d7, d8 = d8,
    d_reliable_cases = d.reliable.cases, d_reliable_hosp = d.reliable.hosp,
    d_hosp_o1 = d.hosp.o1, d_hosp_o2 = d.hosp.o2,
    dhospi = as.numeric(dhospi), n_dhosp = length(dhosp),
    y_dcasei = as.numeric(y.dcasei), o_dcasei = as.numeric(o.dcasei),
    y_dmorti = as.numeric(y.dmorti), o_dmorti = as.numeric(o.dmorti), o_wdmorti = as.numeric(o.wdmorti),
    y_ifr = y.ifr, o_ifr = o.ifr, y_hr = y.hr, o_hr = o.hr, g = g, gtrimp = gtrimp,
    case_nbinom_size1 = case_nbinom_size1, case_nbinom_sizey2 = case_nbinom_sizey2,
    case_nbinom_sizeo2 = case_nbinom_sizeo2, hosp_nbinom_size1 = hosp_nbinom_size1,
    hosp_nbinom_size2 = hosp_nbinom_size2, mort_nbinom_size = mort_nbinom_size, hosp_last7 = hosp_last7)

## handle creation snapshots every global above (EPI_FLAVOUR_AGE2Q = 2)
epi_handle <- .Call("C_epi_create", 2L, as.integer(FitTotalPeriod), as.integer(epi_substeps), epi_globals,
                    as.integer(epi_device))

meta <- .Call("C_epi_df_params", epi_handle)
fit.paramnames <- meta[[1]]
init <- meta[[4]]
df_params <- data.frame(name = fit.paramnames, min = meta[[2]], max = meta[[3]], init = init)
keyparamnames <- c("betay6", "betao6", "betayo6", "betay7", "betao7", "betayo7", "fy9", "fo9", "fyo9", "ifrred")
fitkeyparamnames <- keyparamnames

transformParams <- function(params) params

invTransformParams <- function(posterior) {
    posterior$Tinf <- 1/gamma
    for (k in c("0", "1", "2", "4")) {
        posterior[[paste0("y.Rt", k)]] <- posterior[[paste0("betay", k)]] / gamma
        posterior[[paste0("o.Rt", k)]] <- posterior[[paste0("betao", k)]] / gamma
        posterior[[paste0("yo.Rt", k)]] <- posterior[[paste0("betayo", k)]] / gamma
    }
    posterior
}

## batched: theta is a d x n matrix, one proposal per column
calclogpost_batch <- function(theta) {
    r <- .Call("C_epi_logpost", epi_handle, theta)
    list(logl = r[[1]], logp = r[[2]], status = r[[3]])
}
calclogl_batch <- function(theta) calclogpost_batch(theta)$logl

## the reference's scalar API (drjacoby calls these once per proposal: correct, but the GPU idles)
calclogl <- function(params, x) {
    it <<- it + 1
    calclogpost_batch(matrix(params, ncol = 1))$logl
}
calclogp <- function(params) calclogpost_batch(matrix(params, ncol = 1))$logp

series.names <- c("y.S", "o.S", "y.casei", "o.casei", "y.hospi", "o.hospi", "y.deadi", "o.deadi",
                  "y.E", "y.I", "y.R", "o.E", "o.I", "o.R", "Re", "Rt", "y.Re", "o.Re")
calculateModel_batch <- function(params, period, ext = TRUE) {
    r <- .Call("C_epi_trajectories", epi_handle, params, as.integer(period), ext)
    ns <- if (ext) length(series.names) else 8L
    arr <- array(r[[1]], dim = c(period, ns, ncol(params)))
    list(series = arr, offset = r[[2]], padding = r[[3]])
}
## values of 1..padding as the reference pre-fills them (model-age-groups2q.R:118-141; y.E / o.E are first
## assigned on (padding+1):..., i.e. NA before)
series.prefill <- c(y.N - 1, o.N, 0, 0, 0, 0, 0, 0, NA, 0, 0, NA, 0, 0, 0, 0, 0, 0)
calculateModel <- function(params, period) {
    r <- calculateModel_batch(matrix(params, ncol = 1), period)
    padding <- r$padding[1]
    state <- NULL
    for (k in seq_along(series.names))
        state[[series.names[k]]] <- c(rep(series.prefill[k], padding), r$series[, k, 1])
    state$y.died <- cumsum(state$y.deadi); state$o.died <- cumsum(state$o.deadi)
    state$y.case <- cumsum(state$y.casei); state$o.case <- cumsum(state$o.casei)
    state$padding <- padding
    state$offset <- r$offset[1]
    state
}

## quantileData (libMCMC.R:151-172) on the GPU: `series` is a name from series.names or two names to be summed,
## e.g. quantileData_gpu(data_sample, c("y.hospi", "o.hospi"), 0, 200, c(0.05, 0.5, 0.95)), or "Re" (analyzeMCMC.R:265)
quantileData_gpu <- function(sample, series, startoffset, period, quantiles) {
    params <- apply(sample[, fit.paramnames], 1, function(p) transformParams(unlist(p, use.names = FALSE)))
    idx <- match(series, series.names) - 1L
    .Call("C_epi_quantiles", epi_handle, params, as.integer(period), as.integer(startoffset), idx[1],
          if (length(idx) > 1) idx[2] else -1L, as.numeric(quantiles))
}

## marginalizeData (libMCMC.R:174-188) on the GPU for marginfun = max / sum / min / mean of v[j1:j2], e.g. the
## maximum of Re over the next two months (predictMCMC.R:714-717): marginalizeData_gpu(data_sample, "Re", 0, Range, "max", j1, j2)
marginalizeData_gpu <- function(sample, series, startoffset, period, kind, j1, j2) {
    params <- apply(sample[, fit.paramnames], 1, function(p) transformParams(unlist(p, use.names = FALSE)))
    idx <- match(series, series.names) - 1L
    .Call("C_epi_marginalize", epi_handle, params, as.integer(period), as.integer(startoffset), idx[1],
          if (length(idx) > 1) idx[2] else -1L, match(kind, c("max", "sum", "min", "mean")) - 1L, as.integer(j1), as.integer(j2))
}

calcNominalState <- function(state) {
    state$case <- state$y.case + state$o.case
    state$died <- state$y.died + state$o.died
    state$casei <- state$y.casei + state$o.casei
    state$deadi <- state$y.deadi + state$o.deadi
    state$hospi <- state$y.hospi + state$o.hospi
    state
}
