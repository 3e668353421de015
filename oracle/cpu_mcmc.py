"""CPU Metropolis sampler on the ORACLE density — TEST INFRASTRUCTURE, NOT PRODUCT.

An independent population-adaptive random-walk Metropolis in numpy (proposal covariance 2.38^2/d x pooled covariance,
re-measured during burn-in only; reflection at the df_params box) whose every density evaluation is
oracle.evaluate() on the host threads.  Two users:
  * tests/test_gpu_sampler.py: the statistical reference the GPU samplers must agree with;
  * bench.py's cpu_baseline leg: ESS/s of a CPU sampler on the same target, timed beside the GPU sampler
    (BASELINE.json's metric is "evals/s and ESS/s ... vs R on host CPU"; R is not installed, this is the port).
"""
from __future__ import annotations

import time

import numpy as np


def start_cloud(mn, mx, init, n, seed, width=0.02):
    """start inside the main mode: the target is rugged (HL, DL enter through ceil/floor kernel
    extents, the threshold through an integer data_offset) and has shallow local modes that trap
    ~10 % of widely dispersed starts for any local sampler — which is what the reference's
    20-rung parallel tempering is for (fitMCMC-drj.R:53)"""
    rng = np.random.default_rng(seed)
    return np.clip(init + width * (mx - mn) * (rng.uniform(size=(n, len(init))) - 0.5), mn, mx)


def reflect(y, mn, mx):
    w = mx - mn
    z = np.mod(y - mn, 2 * w)
    z = np.where(z > w, 2 * w - z, z)
    return mn + z


def run(oracle, ds, m, n_chains, burnin, samples, thin, seed, n_threads=16, cov0=None, x0=None, budget_s=None, beta=1.0):
    """Returns (draws[k, chain, d], info).  cov0: proposal covariance to start from (e.g. the posterior covariance
    of a previous run; None: 2 % of the box).  budget_s: stop the SAMPLING phase after that many seconds.
    beta: likelihood tempering exponent (0: the prior, restricted to where the likelihood is finite — what the
    beta = 0 rung of the GPU sampler targets)."""
    rng = np.random.default_rng(seed)
    mn, mx, init = (np.asarray(v) for v in oracle.df_params(ds["flavour"], ds, ds["period"]))
    d = len(init)
    x = start_cloud(mn, mx, init, n_chains, seed + 100) if x0 is None else np.array(x0, dtype=np.float64)

    def post(th):
        r = oracle.evaluate(ds["flavour"], ds, th, ds["period"], m, n_threads=n_threads)
        lp = (beta * r["logl"] if beta != 0 else 0.0) + r["logp"]
        return np.where(np.isfinite(lp) & np.isfinite(r["logl"]) & ((r["status"] & (2 | 16)) == 0), lp, -np.inf)

    lp = post(x)
    evals = n_chains
    if cov0 is None:
        L = np.diag(0.02 * (mx - mn) / np.sqrt(d))
    else:
        L = np.linalg.cholesky(np.asarray(cov0) + 1e-12 * np.diag((mx - mn) ** 2)) * 2.38 / np.sqrt(d)
    keep, accepted, t_samp = [], 0, None
    for it in range(burnin + samples):
        if it == burnin:
            t_samp, evals_samp0 = time.perf_counter(), evals
        if it < burnin and it % 50 == 0 and it > 0 and n_chains > 2 * d:
            C = np.cov(x.T) + 1e-12 * np.diag((mx - mn) ** 2)
            L = np.linalg.cholesky(C) * 2.38 / np.sqrt(d)
        y = reflect(x + rng.standard_normal((n_chains, d)) @ L.T, mn, mx)
        lq = post(y)
        evals += n_chains
        acc = np.log(rng.uniform(size=n_chains)) < lq - lp
        x[acc], lp[acc] = y[acc], lq[acc]
        if it >= burnin:
            accepted += int(acc.sum())
            if (it - burnin) % thin == 0:
                keep.append(x.copy())
            if budget_s is not None and time.perf_counter() - t_samp > budget_s:
                break
    dt = time.perf_counter() - t_samp if t_samp is not None else 0.0
    n_samp_it = max(1, it + 1 - burnin)
    info = dict(sampling_s=dt, sampling_iterations=n_samp_it, evals_sampling=evals - (evals_samp0 if t_samp else evals),
                accept_rate=accepted / (n_samp_it * n_chains))
    return np.array(keep).reshape(len(keep), n_chains, d), info
