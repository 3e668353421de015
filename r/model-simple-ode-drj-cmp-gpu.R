## model-simple-ode-drj-cmp-gpu.R — drop-in for R/models/model-simple-ode-drj-cmp.R (and, with epi_flavour <- 1L,
## for the Ne variant the analyses/*/simple settings name, README.md:78-95).
##
## SOURCE ONLY here (no R in the build image).  Same contract as the age drop-ins: every name the reference file
## defines keeps its meaning; calclogl / calclogp / calculateModel call the GPU through the .Call shim.

dyn.load("epimcmc_shim.so")            # replaces: system("R CMD SHLIB model-simple-ode-cmp.cpp"); dyn.load(...)

InvalidDataOffset <- 10000
Tinc <- 3
died_rate <- 0.007

if (!exists("epi_flavour")) epi_flavour <- 0L         # 0 = simple (8 parameters), 1 = Ne (9: + Nef)
if (!exists("epi_substeps")) epi_substeps <- 4
if (!exists("epi_device")) epi_device <- 0

epi_globals <- list(
    N = N, total_deaths_at_lockdown = total_deaths_at_lockdown, dmort_last = dmort[length(dmort)],
    lockdown_offset = lockdown_offset, lockdown_transition_period = lockdown_transition_period,
    dhospi = as.numeric(dhospi), dmorti = as.numeric(dmorti), n_dhosp = length(dhosp),
    hosp_nbinom_size = hosp_nbinom_size, mort_nbinom_size = mort_nbinom_size)

epi_handle <- .Call("C_epi_create", as.integer(epi_flavour), as.integer(FitTotalPeriod), as.integer(epi_substeps),
                    epi_globals, as.integer(epi_device))

meta <- .Call("C_epi_df_params", epi_handle)
fit.paramnames <- meta[[1]]
init <- meta[[4]]
df_params <- data.frame(name = fit.paramnames, min = meta[[2]], max = meta[[3]], init = init)
keyparamnames <- c("beta0", "betat", "R0", "Rt", "Tinf0", "Tinft")
fitkeyparamnames <- c("R0", "Rt", "Tinf0", "Tinft")
scales <- c(1, 1, 1, 1, 0.05, 1, 1, total_deaths_at_lockdown / 20)   # model-simple-ode-drj-cmp.R:264

## model-simple-ode-drj-cmp.R:136-154
transformParams <- function(params) {
    result <- params
    result[5] <- exp(params[5])
    result
}
invTransformParams <- function(posterior) {
    posterior$Tinc <- Tinc
    posterior$beta0 <- posterior$R0 / posterior$Tinf0
    posterior$betat <- posterior$Rt / posterior$Tinft
    posterior$HR <- exp(posterior$logHR)
    posterior
}

## batched: theta is a d x n matrix of SAMPLER-space vectors, one proposal per column
calclogpost_batch <- function(theta) {
    r <- .Call("C_epi_logpost", epi_handle, theta)
    list(logl = r[[1]], logp = r[[2]], status = r[[3]])
}

## the reference's scalar API: calclogl takes MODEL-space params (the driver passes transformParams(theta),
## fitMCMC-drj.R:19-21), calclogp the untransformed vector (:51)
calclogl <- function(params, x) {
    it <<- it + 1
    theta <- params
    theta[5] <- log(params[5])
    calclogpost_batch(matrix(theta, ncol = 1))$logl
}
calclogp <- function(params) calclogpost_batch(matrix(params, ncol = 1))$logp

series.names <- c("S", "hospi", "deadi", "E", "I", "R", "Re", "Rt")
series.prefill <- c(N - 1, 0, 0, 1, 0, 0, 0, 0)          # model-simple-ode-drj-cmp.R:62-70
calculateModel_batch <- function(params, period, ext = TRUE) {
    r <- .Call("C_epi_trajectories", epi_handle, params, as.integer(period), ext)
    ns <- if (ext) length(series.names) else 3L
    list(series = array(r[[1]], dim = c(period, ns, ncol(params))), offset = r[[2]], padding = r[[3]])
}
calculateModel <- function(params, period) {
    r <- calculateModel_batch(matrix(params, ncol = 1), period)
    padding <- r$padding[1]
    state <- NULL
    for (k in seq_along(series.names))
        state[[series.names[k]]] <- c(rep(series.prefill[k], padding), r$series[, k, 1])
    state$died <- cumsum(state$deadi)
    state$hosp <- cumsum(state$hospi)
    state$padding <- padding
    state$offset <- r$offset[1]
    state
}
calcNominalState <- function(state) state

quantileData_gpu <- function(sample, series, startoffset, period, quantiles) {
    params <- apply(sample[, fit.paramnames], 1, function(p) transformParams(unlist(p, use.names = FALSE)))
    idx <- match(series, series.names) - 1L
    .Call("C_epi_quantiles", epi_handle, params, as.integer(period), as.integer(startoffset), idx[1],
          if (length(idx) > 1) idx[2] else -1L, as.numeric(quantiles))
}
