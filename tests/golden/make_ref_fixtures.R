#!/usr/bin/env Rscript
## make_ref_fixtures.R — mints golden vectors from the UNMODIFIED reference (R + deSolve required).
##
##   Rscript tests/golden/make_ref_fixtures.R <path to the epi-mcmc checkout> [tests/golden]
##
## For every tests/golden/ref_inputs/<dataset>.R (written by make_ref_inputs.py from the committed datasets and
## golden theta) this script
##   1. defines the globals the reference model file reads (data vectors, calendar constants, settings sizes),
##   2. copies the reference's model file and its compiled RHS source into a scratch directory (the model file runs
##      `R CMD SHLIB` beside itself, R/models/model-age-groups2q.R:2-3) and source()s the model file as it is,
##   3. evaluates, for every golden theta, calclogp(theta) and the BODY of calclogl(theta) in an environment it keeps,
##      so that the per-term locals (y.loglC, o.loglC, loglH, loglHRatio, y.loglD, o.loglD) can be read back without
##      touching the reference's code, plus state$offset and state$padding,
##   4. writes tests/golden/ref_<dataset>.csv (one row per theta; deSolve's defaults: lsoda, rtol = atol = 1e-6).
## tests/test_ref_fixtures.py checks the CPU oracle against these files whenever they exist (at LSODA tolerance).
## Nothing here is executed in the build image (no R); it is the committed recipe for pinning the oracle to real R.

args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 1) stop("usage: make_ref_fixtures.R <epi-mcmc checkout> [golden dir]")
ref.root <- normalizePath(args[1])
golden.dir <- normalizePath(if (length(args) > 1) args[2] else file.path("tests", "golden"))
inputs <- list.files(file.path(golden.dir, "ref_inputs"), pattern = "\\.R$", full.names = TRUE)
if (length(inputs) == 0) stop("no ref_inputs/*.R: run tests/golden/make_ref_inputs.py first")

term.names <- c("y.loglC", "o.loglC", "loglH", "loglHRatio", "y.loglD", "o.loglD")

for (inp in inputs) {
    name <- sub("\\.R$", "", basename(inp))
    cat("== ", name, "\n")
    ## a fresh global state per dataset
    rm(list = setdiff(ls(envir = globalenv()),
                      c("args", "ref.root", "golden.dir", "inputs", "inp", "name", "term.names")), envir = globalenv())
    source(inp)                                      # data globals + ref.model, ref.rhs, ref.theta
    scratch <- tempfile(paste0("epi_ref_", name, "_"))
    dir.create(scratch)
    file.copy(file.path(ref.root, "R", "models", c(ref.model, ref.rhs)), scratch)
    old <- setwd(scratch)
    it <- 0                                          # calclogl increments a global counter
    graphs <- function() NULL                        # only called every 1000 evaluations
    source(ref.model)                                # UNMODIFIED: compiles + dyn.loads its RHS, defines the API
    n <- ncol(ref.theta)
    out <- data.frame(i = seq_len(n) - 1, logl = NA_real_, logp = NA_real_, offset = NA_integer_,
                      padding = NA_integer_)
    for (tn in term.names) out[[tn]] <- NA_real_
    for (k in seq_len(n)) {
        p <- ref.theta[, k]
        out$logp[k] <- tryCatch(calclogp(p), error = function(e) NA_real_)
        env <- new.env(parent = globalenv())
        env$params <- transformParams(p)
        env$x <- NULL
        res <- tryCatch(eval(body(calclogl), envir = env), error = function(e) NA_real_)
        out$logl[k] <- res
        st <- get("state", envir = globalenv())      # calclogl assigns it with <<-
        out$offset[k] <- as.integer(st$offset)
        out$padding[k] <- as.integer(st$padding)
        for (tn in term.names)
            if (exists(tn, envir = env, inherits = FALSE)) out[[tn]][k] <- get(tn, envir = env)
    }
    setwd(old)
    dest <- file.path(golden.dir, paste0("ref_", name, ".csv"))
    write.csv(format(out, digits = 17), dest, row.names = FALSE, quote = FALSE)
    cat("wrote ", dest, "\n")
}
cat("R ", R.version.string, ", deSolve ", as.character(packageVersion("deSolve")), "\n", sep = "")
