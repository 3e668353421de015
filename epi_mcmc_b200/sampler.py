"""On-device population sampler (K2) behind the reference driver's call shape.

The reference driver (R/MCMC/fitMCMC-drj.R:47-62) calls drjacoby::run_mcmc four
times, prints ESS and acceptance, and writes the cold-rung sampling draws to
`<outputfile>_<chain>` with the columns `chain, phase, iteration, <params>,
logprior, loglikelihood`.  `run_mcmc` below keeps that call shape and output
schema, but the chains live on the GPU: every iteration is the four K1 launches
plus one fused accept/propose launch (csrc/epi_sampler.cu), the host only
sequences launches and, for DE-MC on several GPUs, triggers the NCCL all-gather
of the population every `exchange_every` iterations (dist.py).
"""
from __future__ import annotations

import csv
import ctypes as C

import numpy as np

from . import _capi

KINDS = {"rwm": _capi.SAMPLER_RWM, "demc": _capi.SAMPLER_DEMC, "drj": _capi.SAMPLER_DRJ}


class Sampler:
    def __init__(self, model, n_chains: int, kind: str = "demc", seed: int = 42, scale: float | None = None,
                 de_eps: float = 1e-4, beta: float = 1.0, chain_id0: int = 0, theta0=None, rungs: int = 1,
                 GTI_pow: float = 1.0, subspace: bool = True, replicates: bool = False, single: bool = True):
        """n_chains is the total on this device; with `rungs` > 1 (parallel tempering, the
        reference driver uses rungs = 20, GTI_pow = 1: fitMCMC-drj.R:53-56) chain i sits on rung
        i // (n_chains // rungs) and the LAST rung is the cold one.

        kind "drj" is drjacoby's own update (one parameter at a time on the logit-reparameterised box,
        per-parameter Robbins-Monro bandwidths towards 0.44 acceptance during burn-in, Metropolis coupling of
        adjacent rungs after every sweep): one `run` iteration is one sweep = d evaluations per chain.

        single=True (default): a quarter of the DE-MC proposals moves one coordinate only (see epi_sampler.cu).

        replicates=True turns the `rungs` into independent replicate populations at beta = 1 (no swaps): the
        spread of their means measures the Monte-Carlo error of the run (`replicate_stats`)."""
        self.model = model
        self._lib = model._lib
        self._h = model._h
        self.n_chains = int(n_chains)
        self.kind = kind
        self.rungs = max(1, int(rungs))
        self.chains_per_rung = self.n_chains // self.rungs
        if scale is None:
            # DE-MC: multiplier of ter Braak's gamma = 2.38/sqrt(2d).  RWM: multiplier of
            # 2.38/sqrt(d) x chol(population covariance) once burn-in adaptation has measured it
            # (before that: of 2 % of the box width / sqrt(d))
            scale = 1.0
        cfg = _capi.SamplerCfg(KINDS[kind], self.n_chains, int(chain_id0), int(seed), float(scale), float(de_eps),
                               float(beta), self.rungs,
                               (0 if subspace else _capi.SAMPLER_FLAG_FULLDIM) | (_capi.SAMPLER_FLAG_REPLICATES if replicates else 0)
                               | (0 if single else _capi.SAMPLER_FLAG_NO_SINGLE),
                               float(GTI_pow))
        # subspace: DE-MC proposals move a random subset of the coordinates (DREAM-style crossover)
        t0 = None
        if theta0 is not None:
            t0 = np.ascontiguousarray(np.asarray(theta0, dtype=np.float64))
            assert t0.shape == (self.n_chains, model.d)
        _capi.check(self._lib.epi_sampler_init(self._h, C.byref(cfg), _capi.as_dp(t0) if t0 is not None else None),
                    self._h)
        self.scale = float(scale)
        self.replicates = bool(replicates)
        self._tune_acc, self._tune_it, self._tune_k = 0, 0, 0

    def run(self, n_iter: int):
        _capi.check(self._lib.epi_sampler_run(self._h, int(n_iter)), self._h)

    def set_scale(self, scale: float):
        _capi.check(self._lib.epi_sampler_set_scale(self._h, float(scale)), self._h)
        self.scale = float(scale)

    def tune_scale(self, comm=None, target: float = 0.2, lo: float = 0.05, hi: float = 20.0):
        """Burn-in only: one Robbins-Monro step of the GLOBAL step multiplier towards `target` acceptance,
        using the acceptances of ALL chains (all ranks when `comm` is given, so the result does not depend on
        the sharding) since the previous call.  A population that starts too tight is accepted too often and
        grows its steps, one that starts too wide shrinks them; the multiplier is frozen afterwards."""
        st = self.stats()
        acc, its = st["accepted"] - self._tune_acc, st["iterations"] - self._tune_it
        self._tune_acc, self._tune_it = st["accepted"], st["iterations"]
        if its <= 0:
            return self.scale
        tot = np.array([float(acc), float(its) * self.n_chains])
        if comm is not None and comm.world > 1:
            tot = comm.allreduce_sum(tot)
        rate = tot[0] / max(1.0, tot[1])
        self._tune_k += 1
        gain = max(0.15, 1.0 / np.sqrt(self._tune_k))
        new = float(np.clip(self.scale * np.exp(gain * (rate - target) / max(target, 1e-3)), lo, hi))
        self.set_scale(new)
        return new

    def adapt(self, enable: bool, target: float = 0.0):
        _capi.check(self._lib.epi_sampler_adapt(self._h, int(enable), float(target)), self._h)

    def record(self, thin: int, capacity: int):
        _capi.check(self._lib.epi_sampler_record(self._h, int(thin), int(capacity)), self._h)

    def refresh_population(self):
        """single-GPU DE-MC: re-snapshot the population from the current states"""
        _capi.check(self._lib.epi_sampler_set_population(self._h, None, 0, 0, 0), self._h)

    def set_population(self, d_pop_ptr: int, n_ranks: int, chains_per_rank: int, ld_rank: int):
        _capi.check(self._lib.epi_sampler_set_population(self._h, d_pop_ptr, n_ranks, chains_per_rank, ld_rank), self._h)

    def state_device(self):
        """(theta_ptr, logl_ptr, logp_ptr, ld): raw device pointers of the d x ld SoA state"""
        th, ll, lp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        ld = C.c_int64()
        _capi.check(self._lib.epi_sampler_state_device(self._h, C.byref(th), C.byref(ll), C.byref(lp), C.byref(ld)), self._h)
        return th.value, ll.value, lp.value, ld.value

    def get(self):
        d, n = self.model.d, self.n_chains
        theta, logl, logp = np.empty((n, d)), np.empty(n), np.empty(n)
        _capi.check(self._lib.epi_sampler_get(self._h, _capi.as_dp(theta), _capi.as_dp(logl), _capi.as_dp(logp)), self._h)
        return theta, logl, logp

    def stats(self) -> dict:
        it, acc, ev = C.c_int64(), C.c_int64(), C.c_int64()
        _capi.check(self._lib.epi_sampler_stats(self._h, C.byref(it), C.byref(acc), C.byref(ev)), self._h)
        n = max(1, it.value * self.n_chains)
        if self.kind == "drj":  # one iteration = one sweep over the free parameters, each its own accept / reject
            dp = self.model.df_params
            n *= max(1, int((np.asarray(dp["max"]) > np.asarray(dp["min"])).sum()))
        return dict(iterations=it.value, accepted=acc.value, evals=ev.value, accept_rate=acc.value / n)

    def pt_stats(self) -> dict:
        prop, acc = C.c_int64(), C.c_int64()
        _capi.check(self._lib.epi_sampler_pt_stats(self._h, C.byref(prop), C.byref(acc)), self._h)
        return dict(proposed=prop.value, accepted=acc.value, swap_rate=acc.value / max(1, prop.value))

    def cold(self, a):
        """slice of a per-chain array (last axis = chain, or axis 1 of draws[k, chain, d]) that
        belongs to the cold rung"""
        lo = (self.rungs - 1) * self.chains_per_rung
        a = np.asarray(a)
        return a[:, lo:lo + self.chains_per_rung] if a.ndim >= 2 else a[lo:lo + self.chains_per_rung]

    def draws_subset(self, first: int, n_sub: int):
        """draws[k, chain, d] and logpost[k, chain] of chains [first, first + n_sub) only"""
        kept = C.c_int32()
        _capi.check(self._lib.epi_sampler_draws_subset(self._h, int(first), int(n_sub), None, None, C.byref(kept)), self._h)
        k = kept.value
        out = np.empty((k, self.model.d, n_sub))
        lp = np.empty((k, n_sub))
        if k:
            _capi.check(self._lib.epi_sampler_draws_subset(self._h, int(first), int(n_sub), _capi.as_dp(out),
                                                           _capi.as_dp(lp), C.byref(kept)), self._h)
        return np.ascontiguousarray(out.transpose(0, 2, 1)), lp

    def bandwidths(self):
        """kind "drj": the current proposal bandwidths (n_chains, d) on the phi scale"""
        bw = np.empty((self.n_chains, self.model.d))
        _capi.check(self._lib.epi_sampler_bandwidths(self._h, _capi.as_dp(bw)), self._h)
        return bw

    def draws_device(self):
        """the recorded draws where they lie: a torch view [kept, d, ld] on the device (no copy) and logpost [kept, ld]"""
        import torch
        dp, lp = C.c_void_p(), C.c_void_p()
        kept, ld = C.c_int32(), C.c_int64()
        _capi.check(self._lib.epi_sampler_draws_device(self._h, C.byref(dp), C.byref(lp), C.byref(kept), C.byref(ld)), self._h)
        if not kept.value:
            return None, None
        from .dist import _DevPtr
        d, k, L = self.model.d, kept.value, ld.value
        dev = f"cuda:{self.model.device}"
        dr = torch.as_tensor(_DevPtr(dp.value, k * d * L), device=dev).view(k, d, L)
        lg = torch.as_tensor(_DevPtr(lp.value, k * L), device=dev).view(k, L)
        return dr, lg

    def replicate_stats(self, params=None, comm=None, last=None):
        """Monte-Carlo error of the recorded phase from INDEPENDENT replicate populations (replicates=True): the
        chains of one replicate share partners, so they are not independent draws, but replicates never interact —
        the variance of the replicate means (each the mean over its chains and recorded iterations) IS the variance
        of that estimator.  ESS of the whole run for a posterior mean = B * var_posterior / var(replicate means),
        B replicates (all ranks together when `comm` is given).  Also split-R-hat over one chain per replicate.
        `last`: use only the first `last` recorded draws (ESS of half the window, to check stationarity).
        Returns dict(ess[p], rhat[p], post_sd[p], n_replicates); everything is reduced on the device."""
        import torch
        dr, _ = self.draws_device()
        if last is not None:
            dr = dr[:int(last)]
        n, cpr, B = self.n_chains, self.chains_per_rung, self.rungs
        idx = torch.as_tensor(list(range(self.model.d)) if params is None else list(params), device=dr.device)
        x = dr[:, idx, :n].reshape(dr.shape[0], len(idx), B, cpr)            # [k, p, B, cpr]
        k = x.shape[0]
        rep_mean = x.mean(dim=(0, 3))                                        # [p, B]
        tot_mean = rep_mean.mean(dim=1)                                      # local
        ss_tot = ((x - tot_mean[None, :, None, None]) ** 2).sum(dim=(0, 2, 3))
        cnt = float(k * B * cpr)
        # one chain per replicate for split R-hat: first chain of each replicate, halves of the recorded phase
        c = x[:, :, :, 0]                                                    # [k, p, B]
        h = k // 2
        halves = torch.stack([c[:h], c[h:2 * h]], dim=0)                     # [2, h, p, B]
        hm = halves.mean(dim=1)                                              # [2, p, B]
        hv = halves.var(dim=1, unbiased=True)                                # [2, p, B]
        part = dict(rep_mean=rep_mean.cpu().numpy(), hm=hm.cpu().numpy(), hv=hv.cpu().numpy(), cnt=cnt,
                    s1=(tot_mean * cnt).cpu().numpy(), s2=ss_tot.cpu().numpy())
        parts = [part]
        if comm is not None and comm.world > 1:
            keys_ = ("rep_mean", "hm", "hv", "s1", "s2")
            gathered = {kk: comm.allgather_np(part[kk]) for kk in keys_}
            cnts = comm.allgather_np(np.array([cnt]))
            parts = [dict({kk: gathered[kk][r] for kk in keys_}, cnt=float(cnts[r][0])) for r in range(comm.world)]
        # DE-MC: replicate b spans the ranks (its chains draw partners from all of them); chains without partners
        # (RWM, DRJ) are independent everywhere, so every rank's replicates count separately
        out = combine_replicate_stats(parts, pooled=self.kind == "demc", half_len=h)
        out["kept"] = k
        return out

    def draws(self):
        """(draws[k, chain, d], logpost[k, chain]) of the recorded thinned states"""
        kept = C.c_int32()
        _capi.check(self._lib.epi_sampler_draws(self._h, None, None, C.byref(kept)), self._h)
        k = kept.value
        out = np.empty((k, self.n_chains, self.model.d))
        lp = np.empty((k, self.n_chains))
        if k:
            _capi.check(self._lib.epi_sampler_draws(self._h, _capi.as_dp(out), _capi.as_dp(lp), C.byref(kept)), self._h)
        return out, lp


# ----------------------------------------------------------------------------- ESS
def combine_replicate_stats(parts, pooled: bool, half_len: int) -> dict:
    """Replicate statistics of a whole run from the per-rank pieces Sampler.replicate_stats reduces on the device.
    Each part: rep_mean[p, B] (mean of replicate b over its chains of that rank and the recorded draws), hm / hv[2, p, B]
    (mean / variance of the two halves of one chain per replicate), cnt (states behind the rank's sums), s1[p] (sum),
    s2[p] (sum of squares about the rank's own mean).  pooled: replicate b of every rank is ONE replicate (DE-MC across
    ranks: equal chain counts, so its mean is the mean of the ranks' means); else the ranks' replicates are distinct.
    ESS of the posterior-mean estimator = B_total * var_posterior / var(replicate means); split R-hat over the
    2 * B half-chains (Gelman et al. 2013)."""
    rep = np.mean([q["rep_mean"] for q in parts], axis=0) if pooled else np.concatenate([q["rep_mean"] for q in parts], axis=1)
    hm = np.concatenate([q["hm"] for q in parts], axis=2)
    hv = np.concatenate([q["hv"] for q in parts], axis=2)
    cnts = np.array([q["cnt"] for q in parts], dtype=float)
    s1 = np.array([q["s1"] for q in parts])
    s2 = np.array([q["s2"] for q in parts])
    gmean = s1.sum(0) / cnts.sum()
    lmean = s1 / cnts[:, None]
    var_post = (s2 + cnts[:, None] * (lmean - gmean) ** 2).sum(0) / (cnts.sum() - 1)
    Bt = rep.shape[1]
    ess = Bt * var_post / np.maximum(rep.var(axis=1, ddof=1), 1e-300)
    m_ = np.concatenate([hm[0], hm[1]], axis=1)   # [p, 2 B]
    W = np.concatenate([hv[0], hv[1]], axis=1).mean(axis=1)
    h = float(half_len)
    rhat = np.sqrt(((h - 1) / h * W + m_.var(axis=1, ddof=1)) / np.maximum(W, 1e-300))
    return dict(ess=ess, rhat=rhat, post_sd=np.sqrt(var_post), n_replicates=Bt)


def _spectrum0_ar(x: np.ndarray) -> float:
    """coda::spectrum0.ar: spectral density at frequency zero from an AR(p) fit (Yule-Walker,
    order by AIC, p <= 10*log10(n)).  This is what coda::effectiveSize — and the ESS drjacoby
    prints (fitMCMC-drj.R:58) — is built on."""
    n = len(x)
    x = x - x.mean()
    v = float(x @ x) / n
    if v <= 0:
        return 0.0
    pmax = max(1, min(n - 1, int(np.floor(10 * np.log10(n)))))
    acov = np.array([float(x[: n - k] @ x[k:]) / n for k in range(pmax + 1)])
    # Levinson-Durbin
    best_aic, best = n * np.log(v), (v, np.zeros(0))
    phi = np.zeros(0)
    err = acov[0]
    for p in range(1, pmax + 1):
        k = (acov[p] - phi @ acov[p - 1:0:-1][: p - 1]) / err if p > 1 else acov[1] / err
        phi = np.concatenate([phi - k * phi[::-1], [k]])
        err = err * (1 - k * k)
        if err <= 0:
            break
        aic = n * np.log(err) + 2 * p
        if aic < best_aic:
            best_aic, best = aic, (err * n / (n - (p + 1)), phi.copy())
    var_pred, coef = best
    return float(var_pred / (1.0 - coef.sum()) ** 2)


def effective_size(chain: np.ndarray) -> float:
    """coda::effectiveSize for one chain of scalar draws"""
    x = np.asarray(chain, dtype=np.float64)
    n = len(x)
    if n < 4 or np.var(x) == 0:
        return 0.0
    s0 = _spectrum0_ar(x)
    return float(min(n, n * np.var(x, ddof=1) / s0)) if s0 > 0 else 0.0


def pooled_ess(draws: np.ndarray, max_chains: int = 256) -> np.ndarray:
    """draws[k, chain, d] -> ESS per parameter, summed over chains (independent chains add;
    coda's convention for mcmc.list).  For very wide populations a seeded subset of
    `max_chains` chains is measured and scaled up."""
    k, n, d = draws.shape
    idx = np.arange(n)
    if n > max_chains:
        idx = np.random.default_rng(0).choice(n, max_chains, replace=False)
    ess = np.zeros(d)
    for p in range(d):
        ess[p] = sum(effective_size(draws[:, c, p]) for c in idx)
    return ess * (n / len(idx))


# ------------------------------------------------------------------------ run_mcmc
def run_mcmc(model, burnin: int = 1000, samples: int = 500, chains: int = 4096, kind: str = "demc", seed: int = 42,
             thin: int = 1, exchange_every: int = 10, comm=None, rungs: int = 1, GTI_pow: float = 1.0, **sampler_kw):
    """The call the reference driver makes (fitMCMC-drj.R:48-57) — burn-in then sampling —
    for a whole population of chains on the GPU.  Returns a dict with `output` (drjacoby
    schema, sampling phase), `diagnostics` (ess per parameter, acceptance) and raw `draws`."""
    rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
    # `chains` = chains per rung, like drjacoby's `chains` x `rungs` layout; draws are the cold rung's
    S = Sampler(model, chains * max(1, rungs), kind=kind, seed=seed, chain_id0=rank * chains * max(1, rungs),
                rungs=rungs, GTI_pow=GTI_pow, **sampler_kw)
    if comm is not None and kind == "demc":
        comm.attach(S)

    def advance(n, refresh):
        done = 0
        while done < n:
            step = min(exchange_every, n - done)
            S.run(step)
            done += step
            if kind == "demc" and refresh:
                comm.exchange(S) if comm is not None else S.refresh_population()

    # Burn-in adapts: the partner snapshot follows the population (all-gathered over the ranks every
    # `exchange_every` iterations) and the proposal shape is re-measured.  For the recorded phase BOTH are frozen:
    # partners then come from a fixed archive of past states, the kernel is a fixed symmetric Metropolis kernel and
    # every chain is a Markov chain with the posterior as its invariant law (a snapshot that keeps following the
    # moving chains is an adaptive scheme that needs a growing archive to be valid: ter Braak & Vrugt 2008).
    S.adapt(True)
    advance(burnin, True)
    S.adapt(False)
    if kind == "demc" and comm is not None:
        comm.exchange(S, wait=True)   # the archive of the sampling phase: the population at the end of burn-in
    keep = max(1, samples // thin)
    S.record(thin, keep)
    advance(samples, False)
    draws, logpost = S.draws()
    if rungs > 1 and len(draws):
        draws, logpost = S.cold(draws), S.cold(logpost)
    st = S.stats()
    ess = pooled_ess(draws) if len(draws) >= 4 else np.zeros(model.d)
    return dict(draws=draws, logpost=logpost, sampler=S,
                diagnostics=dict(ess=dict(zip(model.fit_paramnames, ess)), accept_rate=st["accept_rate"],
                                 evals=st["evals"], iterations=st["iterations"]))


def write_chain_csv(path: str, model, draws: np.ndarray, chain: int, logprior=None, loglike=None, label=None):
    """One chain's sampling draws in the file format fitMCMC-drj.R:60-62 writes with
    write.csv(subset(output, rung == 20 & phase == "sampling")) and libMCMC.R:6-24 reads back:
    a row-name column, then chain, rung, phase, iteration, <fit.paramnames>, logprior, loglikelihood."""
    k = draws.shape[0]
    with open(path, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_NONNUMERIC)
        w.writerow(["", "chain", "rung", "phase", "iteration"] + list(model.fit_paramnames) + ["logprior", "loglikelihood"])
        for i in range(k):
            row = [str(i + 1), label or f"chain{chain + 1}", 20, "sampling", i + 1] + [float(v) for v in draws[i, chain]]
            row += [float(logprior[i, chain]) if logprior is not None else float("nan"),
                    float(loglike[i, chain]) if loglike is not None else float("nan")]
            w.writerow(row)


def write_max_csv(path: str, theta_best: np.ndarray):
    """max.csv warm-start file read by fitMCMC-drj.R:8-13 (`init <- r$x[2:(length(init)+1)]`):
    write.csv of c(maxLogP, params) as column x (model-2s-age-k.R:842-856)."""
    with open(path, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_NONNUMERIC)
        w.writerow(["", "x"])
        for i, v in enumerate(theta_best):
            w.writerow([str(i + 1), float(v)])


def read_max_csv(path: str, d: int) -> np.ndarray:
    """The warm start of fitMCMC-drj.R:8-13: `r <- read.csv("max.csv"); init <- r$x[2:(length(init) + 1)]`
    (x[1] is the log-posterior of the stored maximum)."""
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    col = rows[0].index("x")
    x = [float(r[col]) for r in rows[1:]]
    if len(x) < d + 1:
        raise ValueError(f"{path}: {len(x)} values, need 1 + {d}")
    return np.array(x[1:d + 1])


def read_chain_csv(files, model) -> dict:
    """readData (libMCMC.R:6-24): read and row-bind one or more sample files, then invTransformParams.
    Returns a dict of columns (numeric columns as float arrays)."""
    cols: dict = {}
    for path in ([files] if isinstance(files, str) else list(files)):
        with open(path, newline="") as f:
            rd = csv.reader(f)
            header = next(rd)
            data = list(rd)
        for j, name in enumerate(header):
            cols.setdefault(name or "X", []).extend(r[j] for r in data)
    out = {}
    for name, v in cols.items():
        try:
            out[name] = np.array([float(x) for x in v])
        except ValueError:
            out[name] = np.array(v)
    return model.invTransformParams(out)


def dic(model, posterior: dict) -> dict:
    """DIC() of analyzeMCMC.R:26-44 on a posterior table (columns = fit.paramnames + loglikelihood):
    D.bar = -2 mean(lik), D.hat = -2 calclogl(transformParams(mean theta)) — one evaluation on the GPU."""
    lik = np.asarray(posterior["loglikelihood"], dtype=np.float64)
    theta_bar = np.array([np.mean(posterior[nm]) for nm in model.fit_paramnames])
    d_bar = -2.0 * float(np.mean(lik))
    d_hat = -2.0 * model.calclogl(model.transformParams(theta_bar))
    p_d = d_bar - d_hat
    p_v = float(np.var(-2.0 * lik, ddof=1)) / 2.0
    return dict(DIC=p_d + d_bar, DIC2=p_v + d_bar, IC=2 * p_d + d_bar, pD=p_d, pV=p_v, Dbar=d_bar, Dhat=d_hat)
