## fitMCMC-gpu.R — GPU replacement for R/MCMC/fitMCMC-drj.R.
##
## SOURCE ONLY here (no R in the build image).  Run it from an analysis directory exactly like the
## reference driver (Rscript ../../../r/fitMCMC-gpu.R), with `fitmodel` in settings.R pointing at one of
## the r/model-*-gpu.R drop-ins.  It keeps the reference driver's contract — settings.R / data / model file
## sourced in the same order, warm start from max.csv, one "<outputfile>_<k>" sample file per chain with the
## columns libMCMC.R:6-24 reads back — but the Markov chains run on the device: drjacoby proposes one
## parameter of one chain at a time (a batch of 1 per GPU call); here a population of `gpu_chains` chains per
## rung advances together, one batched log-posterior evaluation per iteration.

source("settings.R")

source(data, chdir=T)
source(fitmodel, chdir=T)

if (file.exists("max.csv")) {
    print("Loading initial state from max.csv")
    r <- read.csv("max.csv")
    init <- r$x[2:(length(init) + 1)]
    df_params$init <- init
}

source(paste(Rdir, "lib/libfit.R", sep=""))

options(scipen=999)

## source it again so that you can override things
source("settings.R")

if (!exists("gpu_chains")) gpu_chains <- 4096     # chains per rung on the device
if (!exists("gpu_files")) gpu_files <- 4          # sample files written (the reference writes 4 chains)
if (!exists("gpu_kind")) gpu_kind <- 1L           # 1 = DE-MC, 0 = adaptive random-walk Metropolis
if (!exists("gpu_rungs")) gpu_rungs <- 1L         # drjacoby's rungs = 20 is available; DE-MC does not need it
if (!exists("gpu_seed")) gpu_seed <- 42

print(fit.paramnames)
r0 <- calclogpost_batch(matrix(init, ncol = 1))
print(c("initial logl: ", r0$logl, " logl prior: ", r0$logp))

## start every chain at init + 2 % box jitter (NULL lets the library do exactly that), or warm-start them all
## around the max.csv point the same way
out <- .Call("C_epi_run_mcmc", epi_handle, NULL, as.integer(gpu_chains), as.integer(1e3), as.integer(0.5e3), 1L,
             as.integer(gpu_rungs), 1.0, as.numeric(gpu_seed), as.integer(gpu_kind))
draws <- out[[1]]                 # d x chains x kept
logpost <- out[[2]]               # chains x kept
print(c("acceptance rate: ", out[[3]], " evaluations: ", out[[4]]))

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
