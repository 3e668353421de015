"""GPU tests of K2 (fused propose/accept) and the on-device sampler."""
import ctypes as C

import numpy as np
import pytest

import philox_ref

pytestmark = pytest.mark.gpu


def _model(name="S120", m=2):
    from epi_mcmc_b200.model import Model, load_dataset
    return Model(load_dataset(name), substeps=m)


def test_philox_device_matches_host_reference_and_kat():
    # Random123 known-answer vectors for philox4x32-10
    assert philox_ref.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox_ref.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    M = _model()
    n = 257
    out = np.zeros(4 * n, dtype=np.uint32)
    seed, it, use = 0x0123456789ABCDEF, 77, 3
    rc = M._lib.epi_philox_probe(M._h, C.c_uint64(seed), it, use, n, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == 0
    for j in (0, 1, 31, 256):
        assert tuple(int(v) for v in out[4 * j:4 * j + 4]) == philox_ref.philox4x32_10((j, 0, it, use), (seed & 0xFFFFFFFF, seed >> 32))
    M.close()


def test_deterministic_and_sharding_invariant():
    """same (seed, chain id, iteration) -> same stream: a run is reproducible, and chains
    [0, n) on one GPU equal chains [0, n/2) + [n/2, n) run as two shards (RWM has no coupling)."""
    from epi_mcmc_b200.sampler import Sampler
    n = 256
    M = _model()
    a = Sampler(M, n, kind="rwm", seed=5)
    a.run(15)
    ta, la, pa = a.get()
    b = Sampler(M, n, kind="rwm", seed=5)
    b.run(15)
    tb, lb, pb = b.get()
    assert np.array_equal(ta, tb) and np.array_equal(la, lb)
    lo = Sampler(M, n // 2, kind="rwm", seed=5, chain_id0=0)
    lo.run(15)
    t0, _, _ = lo.get()
    hi = Sampler(M, n // 2, kind="rwm", seed=5, chain_id0=n // 2)
    hi.run(15)
    t1, _, _ = hi.get()
    assert np.array_equal(np.vstack([t0, t1]), ta)
    c = Sampler(M, n, kind="rwm", seed=6)
    c.run(15)
    assert not np.array_equal(c.get()[0], ta)
    M.close()


def test_state_is_consistent_and_inside_the_box():
    """the stored log densities are those of the stored states (re-evaluated through epi_eval),
    every state is inside df_params, counters add up"""
    from epi_mcmc_b200.sampler import Sampler
    M = _model("A300", 4)
    S = Sampler(M, 384, kind="demc", seed=9)
    for _ in range(3):
        S.run(10)
        S.refresh_population()
    th, ll, lp = S.get()
    l2, p2, st = M.calclogpost_batch(th)
    assert np.array_equal(ll, l2) and np.array_equal(lp, p2)
    assert (th >= M.df_params["min"]).all() and (th <= M.df_params["max"]).all()
    s = S.stats()
    assert s["iterations"] == 30 and s["evals"] == 384 * 31 and 0 < s["accepted"] < 384 * 30
    M.close()


def _start(mn, mx, init, n, seed):
    from oracle import cpu_mcmc
    return cpu_mcmc.start_cloud(mn, mx, init, n, seed)


def _cpu_reference_sampler(oracle, ds, m, n_chains, burnin, samples, thin, seed):
    """An independent CPU sampler on the ORACLE density (oracle/cpu_mcmc.py): population-adaptive Metropolis in
    numpy (proposal covariance = 2.38^2/d x pooled covariance, refreshed during burn-in only)."""
    from oracle import cpu_mcmc
    draws, _ = cpu_mcmc.run(oracle, ds, m, n_chains, burnin, samples, thin, seed)
    return draws.reshape(-1, draws.shape[-1])


def test_samplers_agree_on_the_posterior(oracle):
    """One target (simple model, synthetic S120), three samplers: GPU DE-MC, GPU adaptive-covariance
    RWM, and an independent CPU Metropolis sampler on the oracle density.  Posterior means and
    10/90 % quantiles agree within Monte-Carlo error (the north_star's statistical parity check;
    the reference's own sampler, drjacoby, is not installed)."""
    from epi_mcmc_b200.model import load_dataset
    from epi_mcmc_b200.sampler import run_mcmc
    ds = load_dataset("S120")
    M = _model("S120", 1)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    r1 = run_mcmc(M, burnin=8000, samples=4000, chains=2048, kind="demc", seed=1, thin=40)
    r2 = run_mcmc(M, burnin=8000, samples=4000, chains=2048, kind="rwm", seed=2, thin=40)
    d1 = r1["draws"].reshape(-1, M.d)
    d2 = r2["draws"].reshape(-1, M.d)
    d3 = _cpu_reference_sampler(oracle, ds, 1, n_chains=256, burnin=4000, samples=3000, thin=10, seed=3)
    sd = d3.std(axis=0)
    print("means", d1.mean(0), d2.mean(0), d3.mean(0), "sds", d1.std(0), d2.std(0), sd)
    for name, dd in (("demc", d1), ("rwm", d2)):
        z = np.abs(dd.mean(axis=0) - d3.mean(axis=0)) / sd
        assert (z < 0.25).all(), (name, z)
        qa, qb = np.quantile(dd, [0.1, 0.9], axis=0), np.quantile(d3, [0.1, 0.9], axis=0)
        assert (np.abs(qa - qb) / sd < 0.3).all(), (name, np.abs(qa - qb) / sd)
        assert (np.abs(dd.std(axis=0) / sd - 1) < 0.35).all(), (name, dd.std(axis=0) / sd)
    truth = M.df_params["init"]
    lo, hi = np.quantile(d1, [0.001, 0.999], axis=0)
    assert ((truth >= lo) & (truth <= hi)).all()
    print("accept", r1["diagnostics"]["accept_rate"], r2["diagnostics"]["accept_rate"], "ess", r1["diagnostics"]["ess"], r2["diagnostics"]["ess"])
    assert 0.03 < r1["diagnostics"]["accept_rate"] < 0.8 and 0.03 < r2["diagnostics"]["accept_rate"] < 0.8
    assert min(r1["diagnostics"]["ess"].values()) > 200
    M.close()


def test_parallel_tempering_prior_rung_and_cold_rung(oracle):
    """Rungs as in the reference driver (beta_r = (r/(R-1))^GTI_pow, adjacent-rung swaps).  The
    beta = 0 rung samples the PRIOR restricted to the df_params box, which is known in closed
    form — an analytic check of propose/accept/reflect/swap; the cold rung reproduces the posterior
    of the un-tempered sampler."""
    from scipy import stats
    from epi_mcmc_b200.sampler import Sampler, run_mcmc
    M = _model("S120", 1)
    mn, mx = M.df_params["min"], M.df_params["max"]
    R, cpr = 6, 512
    S = Sampler(M, R * cpr, kind="demc", seed=4, rungs=R, GTI_pow=2.0)
    S.adapt(True)
    for _ in range(300):
        S.run(10)
        S.refresh_population()
    S.adapt(False)
    S.record(20, 100)
    for _ in range(200):
        S.run(10)
        S.refresh_population()
    draws, _ = S.draws()
    hot = draws[:, :cpr, :].reshape(-1, M.d)          # rung 0: beta = 0
    # uniform components of the prior (R0, Rt, logHR, HL, lockdownmort): mean = mid-box, sd = width/sqrt(12)
    for p in (0, 1, 4, 5, 7):
        w = mx[p] - mn[p]
        assert abs(hot[:, p].mean() - 0.5 * (mn[p] + mx[p])) < 0.03 * w, p
        assert abs(hot[:, p].std() - w / np.sqrt(12)) < 0.03 * w, p
    # DL ~ N(21, 4) truncated to its box (model-simple-ode-drj-cmp.R:174)
    tn = stats.truncnorm((mn[6] - 21) / 4, (mx[6] - 21) / 4, loc=21, scale=4)
    assert abs(hot[:, 6].mean() - tn.mean()) < 0.15 and abs(hot[:, 6].std() - tn.std()) < 0.15
    # G0 = 3 + Tinf0/2 ~ N(5, 1) and G0 - Gt ~ N(0, 1.5) (:175-176): Tinf0 - Tinft ~ N(0, 3) before truncation
    assert abs(hot[:, 2].mean() - 4.0) < 0.25
    st = S.pt_stats()
    assert st["proposed"] > 0 and 0.02 < st["swap_rate"] < 0.99
    cold = S.cold(draws).reshape(-1, M.d)
    plain = run_mcmc(M, burnin=5000, samples=2000, chains=1024, kind="demc", seed=5, thin=20)["draws"].reshape(-1, M.d)
    z = np.abs(cold.mean(axis=0) - plain.mean(axis=0)) / plain.std(axis=0)
    assert (z < 0.35).all(), z
    assert (hot.std(axis=0) > 1.5 * cold.std(axis=0))[[0, 1, 4]].all()
    M.close()


def test_chain_csv_schema(tmp_path):
    from epi_mcmc_b200.sampler import run_mcmc, write_chain_csv
    import csv
    M = _model("S120", 1)
    r = run_mcmc(M, burnin=20, samples=10, chains=64, kind="rwm", seed=3)
    path = tmp_path / "run_1"
    write_chain_csv(str(path), M, r["draws"], 0)
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["", "chain", "rung", "phase", "iteration"] + M.fit_paramnames + ["logprior", "loglikelihood"]
    assert len(rows) == 11 and rows[1][3] == "sampling"
    M.close()


def test_sample_file_round_trip_warm_start_and_dic(tmp_path, oracle):
    """The files either side of the path: write.csv schema (fitMCMC-drj.R:60-62) -> readData
    (libMCMC.R:6-24, incl. invTransformParams) -> DIC (analyzeMCMC.R:26-44); max.csv -> warm start
    (fitMCMC-drj.R:8-13)."""
    from epi_mcmc_b200.model import load_dataset
    from epi_mcmc_b200.sampler import (dic, read_chain_csv, read_max_csv, run_mcmc, write_chain_csv, write_max_csv)
    M = _model("S120", 2)
    r = run_mcmc(M, burnin=40, samples=30, chains=64, kind="demc", seed=5)
    draws = r["draws"]
    k, n, d = draws.shape
    flat = draws.reshape(k * n, d)
    ll, lp, _ = M.calclogpost_batch(flat)
    ll, lp = ll.reshape(k, n), lp.reshape(k, n)
    files = []
    for c in (0, 1):
        path = tmp_path / f"run_{c + 1}"
        write_chain_csv(str(path), M, draws, c, logprior=lp, loglike=ll)
        files.append(str(path))
    post = read_chain_csv(files, M)
    assert len(post["loglikelihood"]) == 2 * k
    for j, nm in enumerate(M.fit_paramnames):
        assert np.array_equal(post[nm], np.concatenate([draws[:, 0, j], draws[:, 1, j]]))
    assert np.allclose(post["beta0"], post["R0"] / post["Tinf0"]) and np.allclose(post["HR"], np.exp(post["logHR"]))
    got = dic(M, post)
    lik = post["loglikelihood"]
    theta_bar = np.array([post[nm].mean() for nm in M.fit_paramnames])
    ref_hat = oracle.evaluate("simple", load_dataset("S120"), theta_bar[None], 120, 2)["logl"][0]
    d_bar = -2 * lik.mean()
    assert abs(got["Dbar"] - d_bar) < 1e-9 and abs(got["Dhat"] + 2 * ref_hat) <= 1e-9 * abs(ref_hat)
    assert abs(got["DIC"] - (2 * d_bar + 2 * ref_hat)) <= 1e-8 * abs(d_bar)
    assert abs(got["pV"] - np.var(-2 * lik, ddof=1) / 2) < 1e-9
    # warm start: max.csv = c(maxLogP, params)
    best = np.unravel_index(np.argmax(ll + lp), ll.shape)
    write_max_csv(str(tmp_path / "max.csv"), np.concatenate([[ll[best] + lp[best]], draws[best]]))
    assert np.array_equal(read_max_csv(str(tmp_path / "max.csv"), M.d), draws[best])
    M.close()


def test_sampler_runs_on_the_secondary_age_model():
    """K2 is flavour-agnostic: DE-MC on the 36-parameter age-2i model moves, stays inside the box and its
    cached log densities are those of a fresh evaluation of the states."""
    from epi_mcmc_b200.sampler import Sampler
    M = _model("I290", 2)
    S = Sampler(M, 1024, kind="demc", seed=21)
    for _ in range(4):
        S.run(10)
        S.refresh_population()
    th, ll, lp = S.get()
    assert th.shape == (1024, 36)
    assert (th >= M.df_params["min"]).all() and (th <= M.df_params["max"]).all()
    l2, p2, st = M.calclogpost_batch(th)
    assert np.array_equal(ll, l2) and np.array_equal(lp, p2)
    assert 0 < S.stats()["accepted"] < 1024 * 40
    M.close()


def test_global_scale_control():
    """epi_sampler_set_scale / Sampler.tune_scale: a smaller global step is accepted more often, the burn-in tuner
    moves the population acceptance rate towards its target, and a run is reproducible through a scale change."""
    from epi_mcmc_b200.sampler import Sampler
    M = _model("S120", 2)

    def rate(S, iters=60):
        a0, i0 = S.stats()["accepted"], S.stats()["iterations"]
        for _ in range(iters // 10):
            S.run(10); S.refresh_population()
        st = S.stats()
        return (st["accepted"] - a0) / ((st["iterations"] - i0) * S.n_chains)

    S = Sampler(M, 2048, kind="demc", seed=5)
    for _ in range(10):
        S.run(10); S.refresh_population()
    r1 = rate(S)
    S.set_scale(0.2)
    r2 = rate(S)
    assert r2 > r1 + 0.05, (r1, r2)
    for _ in range(30):
        S.run(10); S.refresh_population(); S.tune_scale(target=0.3)
    r3 = rate(S)
    assert abs(r3 - 0.3) < 0.08, (r3, S.scale)
    # reproducible: the same sequence of calls gives the same state
    out = []
    for _ in range(2):
        T = Sampler(M, 512, kind="demc", seed=8)
        T.run(10); T.refresh_population(); T.set_scale(0.5); T.run(10)
        out.append(T.get()[0])
    assert np.array_equal(out[0], out[1])
    M.close()


def test_the_r_shim_call_sequence():
    """The exact C-ABI sequence r/epimcmc_shim.c::C_epi_run_mcmc issues (the shim cannot be built here: no R) —
    init from a zeroed cfg with the shim's defaults, burn-in in chunks of 10 with the single-rank population
    refresh, record, sampling, draws of the cold rung by epi_sampler_draws_subset — driven through ctypes."""
    import ctypes as C
    from epi_mcmc_b200 import _capi
    from epi_mcmc_b200.model import Model, load_dataset
    M = Model(load_dataset("S120"), substeps=2)
    lib, h, d = M._lib, M._h, M.d
    cpr, rungs, nburn, nsamp, thin = 96, 2, 35, 45, 3
    n, keep = cpr * rungs, nsamp // thin
    cfg = _capi.SamplerCfg()                                   # memset(&cfg, 0, sizeof cfg)
    cfg.kind, cfg.n_chains, cfg.seed, cfg.scale, cfg.de_eps, cfg.beta = 1, n, 42, 1.0, 1e-4, 1.0
    cfg.n_rungs, cfg.gti_pow = rungs, 1.0
    ck = lambda rc: _capi.check(rc, h)
    ck(lib.epi_sampler_init(h, C.byref(cfg), None))
    ck(lib.epi_sampler_adapt(h, 1, C.c_double(0.0)))
    for phase, total in (("burnin", nburn), ("sampling", nsamp)):
        if phase == "sampling":
            ck(lib.epi_sampler_adapt(h, 0, C.c_double(0.0)))
            ck(lib.epi_sampler_record(h, thin, keep))
        done = 0
        while done < total:
            ck(lib.epi_sampler_run(h, min(10, total - done)))
            ck(lib.epi_sampler_set_population(h, None, 1, n, 0))
            done += 10
    kept = C.c_int32()
    dr, lp = np.empty((keep, d, cpr)), np.empty((keep, cpr))
    ck(lib.epi_sampler_draws_subset(h, n - cpr, cpr, _capi.as_dp(dr), _capi.as_dp(lp), C.byref(kept)))
    it, acc, ev = C.c_int64(), C.c_int64(), C.c_int64()
    ck(lib.epi_sampler_stats(h, C.byref(it), C.byref(acc), C.byref(ev)))
    assert kept.value == keep and it.value == nburn + nsamp and ev.value == (nburn + nsamp + 1) * n
    mn, mx = M.df_params["min"], M.df_params["max"]
    assert np.isfinite(dr).all() and (dr >= mn[None, :, None]).all() and (dr <= mx[None, :, None]).all()
    # logpost of a recorded draw == one batched evaluation of it (cold rung: beta = 1)
    last = np.ascontiguousarray(dr[-1].T)                      # cpr x d
    l, p, s = M.calclogpost_batch(last)
    assert np.array_equal(l + p, lp[-1])
    M.close()
