## fitMCMC-gpu.R — GPU replacement for R/MCMC/fitMCMC-drj.R.
##
## SOURCE ONLY here (no R in the build image).  Run it from an analysis directory exactly like the
## reference driver (Rscript ../../../r/fitMCMC-gpu.R), with `fitmodel` in settings.R pointing at one of
## the r/model-*-gpu.R drop-ins.  It keeps the reference driver's contract — settings.R / data / model file
## sourced in the same order, warm start from max.csv, one "<outputfile>_<k>" sample file per chain with the
## columns libMCMC.R:6-24 reads back — but the Markov chains run on the device: drjacoby proposes one
## parameter of one chain at a time (a batch of 1 per GPU call); here a population of `gpu_chains` chains per
## rung advances together, one batched log-posterior evaluation per iteration.

## the same sequence the reference driver runs before it samples (R/MCMC/fitMCMC-drj.R:3-24): settings, data file,
## model file, the optional warm start, libfit.R, then settings.R a second time so that it can override anything
## the model file set
source("settings.R")
source(data, chdir = TRUE)
source(fitmodel, chdir = TRUE)

warm_start <- file.exists("max.csv")
if (warm_start) {
    ## max.csv: x[1] is the log-posterior of the best point of the previous run, x[2..] the point itself
    best_prev <- read.csv("max.csv")$x
    init <- best_prev[1 + seq_along(init)]
    df_params$init <- init
    print("initial state taken from max.csv")
}

source(paste(Rdir, "lib/libfit.R", sep = ""))
options(scipen = 999)
source("settings.R")

if (!exists("gpu_chains")) gpu_chains <- 4096     # chains per rung on the device
if (!exists("gpu_files")) gpu_files <- 4          # sample files written (the reference writes 4 chains)
if (!exists("gpu_kind")) gpu_kind <- 1L           # 1 = DE-MC, 0 = adaptive random-walk Metropolis, 2 = drjacoby's own
                                                  # one-parameter update (burn-in / samples then count sweeps, as in drjacoby;
                                                  # use it with gpu_rungs <- 20L for the reference's tempering ladder)
if (!exists("gpu_rungs")) gpu_rungs <- 1L         # drjacoby's rungs = 20 is available; DE-MC does not need it
if (!exists("gpu_seed")) gpu_seed <- 42

print(fit.paramnames)
r0 <- calclogpost_batch(matrix(init, ncol = 1))
print(c("initial logl: ", r0$logl, " logl prior: ", r0$logp))

## every chain starts at `init` — the model file's, or the max.csv point loaded above — plus 2 % of the box width of
## jitter, clamped to the df_params box (the handle only knows the model file's compiled-in init, so the starting
## cloud is built here and passed in: a warm start really moves the whole population)
n_start <- as.integer(gpu_chains) * max(1L, as.integer(gpu_rungs))
set.seed(gpu_seed)
box_w <- df_params$max - df_params$min
theta0 <- matrix(init, nrow = length(init), ncol = n_start) +
    0.02 * box_w * (matrix(runif(length(init) * n_start), nrow = length(init)) - 0.5)
theta0 <- pmin(pmax(theta0, df_params$min), df_params$max)       # (recycled down the columns: one bound per row)
storage.mode(theta0) <- "double"
out <- .Call("C_epi_run_mcmc", epi_handle, theta0, as.integer(gpu_chains), as.integer(1e3), as.integer(0.5e3), 1L,
             as.integer(gpu_rungs), 1.0, as.numeric(gpu_seed), as.integer(gpu_kind))
draws <- out[[1]]                 # d x chains x kept
logpost <- out[[2]]               # chains x kept
print(c("acceptance rate (sampling phase): ", out[[3]], " evaluations: ", out[[4]]))

## effective sample size per parameter, pooled over a subset of chains (what drjacoby prints, :58)
if (requireNamespace("coda", quietly = TRUE)) {
    sub <- seq_len(min(64, dim(draws)[2]))
    ess <- sapply(seq_along(fit.paramnames), function(p)
        sum(sapply(sub, function(c) coda::effectiveSize(draws[p, c, ]))) * dim(draws)[2] / length(sub))
    names(ess) <- fit.paramnames
    print(ess)
}

## one sample file per written chain, same columns as
##   write.csv(subset(r_mcmc_out$output, rung == 20 & phase == "sampling"))      (fitMCMC-drj.R:60-62)
## logprior / loglikelihood: recomputed for the written chains with one batched call each
kept <- dim(draws)[3]
for (chain in 1:gpu_files) {
    th <- matrix(draws[, chain, ], nrow = length(fit.paramnames))       # d x kept
    lp <- calclogpost_batch(th)
    output <- data.frame(chain = paste("chain", chain, sep = ''), rung = 20, phase = "sampling",
                         iteration = seq_len(kept), t(th), logprior = lp$logp, loglikelihood = lp$logl)
    names(output)[5:(4 + length(fit.paramnames))] <- fit.paramnames
    write.csv(output, file = paste(outputfile, "_", chain, sep = ''))
}

## the best point seen, in the max.csv layout the warm start above reads (x[1] = its log-posterior)
best <- which(logpost == max(logpost), arr.ind = TRUE)[1, ]
write.csv(data.frame(x = c(max(logpost), draws[, best[1], best[2]])), file = "max.csv")
