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


def test_drjacoby_style_sampler(oracle):
    """EPI_SAMPLER_DRJ, the update of the sampler the reference driver calls (fitMCMC-drj.R:48-57): one parameter
    at a time on the logit-reparameterised box, Robbins-Monro bandwidths towards 0.44 acceptance in burn-in,
    rungs with Metropolis coupling.  Checked: the adapted acceptance sits at the target; the cached densities are
    those of the states; the kernel leaves the posterior INVARIANT (chains started from posterior draws of the DE-MC
    sampler keep their law — one-parameter updates mix slowly along the R0/Tinf ridge, so convergence from a point
    start is not what a short test can show); the beta = 0 rung reproduces the analytic prior (which it only does
    when the Jacobian term of the reparameterisation is right); the cold rung of a tempered run keeps the posterior."""
    from scipy import stats
    from epi_mcmc_b200.sampler import Sampler, run_mcmc
    M = _model("S120", 1)
    mn, mx = M.df_params["min"], M.df_params["max"]
    S = Sampler(M, 1024, kind="drj", seed=11)
    S.adapt(True)
    S.run(300)
    a0 = S.stats()
    S.run(100)
    a1 = S.stats()
    rate = (a1["accepted"] - a0["accepted"]) / (100 * 1024 * M.d)
    assert abs(rate - 0.44) < 0.05, rate
    assert a1["evals"] == 1024 * (1 + 400 * M.d) and a1["iterations"] == 400
    bw = S.bandwidths()
    assert np.isfinite(bw).all() and (bw > 0).all() and bw.std(axis=0).max() > 0
    th, ll, lpr = S.get()
    l2, p2, _ = M.calclogpost_batch(th)
    assert np.array_equal(ll, l2) and np.array_equal(lpr, p2)
    # invariance of the posterior
    d_de = run_mcmc(M, burnin=8000, samples=4000, chains=2048, kind="demc", seed=1, thin=40)["draws"]
    flat = d_de.reshape(-1, M.d)
    sd = flat.std(axis=0)
    n = 2048
    T = Sampler(M, n, kind="drj", seed=12, theta0=d_de[-1])
    T.adapt(True); T.run(150); T.adapt(False)
    T.record(10, 30); T.run(300)
    dr, _ = T.draws()
    assert dr.shape == (30, n, M.d) and (dr > mn).all() and (dr < mx).all()
    d_drj = dr.reshape(-1, M.d)
    z = np.abs(d_drj.mean(axis=0) - flat.mean(axis=0)) / sd
    assert (z < 0.15).all(), z
    assert (np.abs(d_drj.std(axis=0) / sd - 1) < 0.15).all(), d_drj.std(axis=0) / sd
    qa, qb = np.quantile(d_drj, [0.1, 0.9], axis=0), np.quantile(flat, [0.1, 0.9], axis=0)
    assert (np.abs(qa - qb) / sd < 0.2).all(), np.abs(qa - qb) / sd
    # tempered as the driver runs it: every rung starts from posterior draws; the beta = 0 rung relaxes to the prior
    R, cpr = 5, 512
    th0 = np.concatenate([d_de[-1 - r][:cpr] for r in range(R)])
    P = Sampler(M, R * cpr, kind="drj", seed=13, rungs=R, GTI_pow=2.0, theta0=th0)
    P.adapt(True); P.run(400); P.adapt(False)
    P.record(10, 40); P.run(400)
    dt, _ = P.draws()
    cold, hot = P.cold(dt).reshape(-1, M.d), dt[:, :cpr, :].reshape(-1, M.d)
    z = np.abs(cold.mean(axis=0) - flat.mean(axis=0)) / sd
    assert (z < 0.25).all(), z
    for p in (0, 1, 4, 5, 7):   # uniform prior components: mid-box mean, sd = width / sqrt(12)
        w = mx[p] - mn[p]
        assert abs(hot[:, p].mean() - 0.5 * (mn[p] + mx[p])) < 0.03 * w, p
        assert abs(hot[:, p].std() - w / np.sqrt(12)) < 0.03 * w, p
    tn = stats.truncnorm((mn[6] - 21) / 4, (mx[6] - 21) / 4, loc=21, scale=4)   # DL ~ N(21, 4) on its box
    assert abs(hot[:, 6].mean() - tn.mean()) < 0.2 and abs(hot[:, 6].std() - tn.std()) < 0.2
    st = P.pt_stats()
    assert st["proposed"] > 0 and 0.02 < st["swap_rate"] < 0.99
    M.close()


def test_replicate_populations_and_their_monte_carlo_error():
    """EPI_SAMPLER_FLAG_REPLICATES: the rungs are independent populations at beta = 1.  Their means scatter like
    independent estimates (ESS from the spread of the replicate means is finite, positive and below the number of
    recorded states), split R-hat over one chain per replicate is near 1, and two replicates never share partners:
    moving one replicate's start does not change another replicate's trajectory."""
    from epi_mcmc_b200.sampler import Sampler
    M = _model("S120", 1)
    B, cpr = 16, 256
    S = Sampler(M, B * cpr, kind="demc", seed=3, rungs=B, replicates=True)
    S.adapt(True)
    for _ in range(300):
        S.run(10); S.refresh_population()
    S.adapt(False)
    S.record(5, 200)
    S.run(1000)
    r = S.replicate_stats()
    n_states = 200 * B * cpr
    assert r["n_replicates"] == B and r["kept"] == 200
    assert np.isfinite(r["ess"]).all() and (r["ess"] > 50).all() and (r["ess"] < 20 * n_states).all(), r["ess"]
    # split R-hat of SINGLE chains over a 1,000-iteration window: finite and >= ~1; it approaches 1 only over windows of
    # many integrated autocorrelation times (bench.py reports it for its longer window)
    assert np.isfinite(r["rhat"]).all() and (r["rhat"] > 0.95).all() and (r["rhat"] < 6).all(), r["rhat"]
    half = S.replicate_stats(last=100)
    assert half["kept"] == 100 and np.isfinite(half["ess"]).all()
    assert S.pt_stats()["proposed"] == 0
    # independence: same seed, replicate 0 started elsewhere -> replicates 1.. are bit-identical
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(0)
    th0 = np.clip(init + 0.02 * (mx - mn) * (rng.uniform(size=(4 * 128, M.d)) - 0.5), mn, mx)
    outs = []
    for shift in (0.0, 0.01):
        t = th0.copy()
        t[:128] = np.clip(t[:128] + shift * (mx - mn), mn, mx)
        T = Sampler(M, 4 * 128, kind="demc", seed=7, rungs=4, replicates=True, theta0=t)
        T.run(10); T.refresh_population(); T.run(10)
        outs.append(T.get()[0])
    assert np.array_equal(outs[0][128:], outs[1][128:]) and not np.array_equal(outs[0][:128], outs[1][:128])
    M.close()


def test_k2_alone_and_init_validation():
    """epi_k2_bench times the fused accept/propose kernel on a synthetic population (no K1 workspace); a sampler
    configuration that fails validation leaves no half-built sampler behind."""
    import ctypes as C
    from epi_mcmc_b200 import _capi
    M = _model("A300", 4)
    us, by, ar = C.c_double(), C.c_double(), C.c_double()
    for kind in (_capi.SAMPLER_DEMC, _capi.SAMPLER_DRJ):
        _capi.check(M._lib.epi_k2_bench(M._h, kind, 1 << 16, 20, C.byref(us), C.byref(by), C.byref(ar)), M._h)
        assert us.value > 0 and by.value > 0, (kind, us.value, by.value, ar.value)
        if kind == _capi.SAMPLER_DEMC:   # a fixed quarter of the chains accepts every proposal, the others none
            assert 0.23 < ar.value < 0.27, ar.value
    cfg = _capi.SamplerCfg()
    cfg.kind, cfg.n_chains, cfg.seed, cfg.scale, cfg.de_eps, cfg.beta, cfg.n_rungs, cfg.gti_pow = 1, 100, 1, 1.0, 1e-4, 1.0, 3, 1.0
    assert M._lib.epi_sampler_init(M._h, C.byref(cfg), None) == 1        # EPI_ERR_ARG: 100 % 3 != 0
    assert M._lib.epi_sampler_run(M._h, 1) == 5                           # EPI_ERR_STATE: nothing was installed
    cfg.n_chains, cfg.n_rungs = 4, 2
    assert M._lib.epi_sampler_init(M._h, C.byref(cfg), None) == 1        # DE-MC needs 3 chains per rung
    assert M._lib.epi_sampler_run(M._h, 1) == 5
    M.close()


def _robust(d):
    q = np.quantile(d, [0.25, 0.5, 0.75], axis=0)
    return q[1], (q[2] - q[0]) / 1.349


def test_age_model_statistical_parity(oracle):
    """Statistical check of K2 on the AGE model (45 parameters, 38 coupled prior terms), against an independent CPU
    Metropolis sampler on the oracle density, where a test can afford convergence: the beta = 0 rung of a tempered
    run targets the prior restricted to the box and to a finite likelihood — smooth, so both samplers mix in
    seconds — and all 45 marginals must agree.  (On the age POSTERIOR neither sampler traverses the law within a
    test's budget: the integrated autocorrelation time of a single DE-MC chain is beyond 10^4 iterations — bench.py's
    split R-hat — and the CPU random walk accepts < 1 % of its proposals; there the check is GPU <-> oracle parity of
    the density to 1e-10 plus the sampler checks on the simple model and on this rung.)"""
    import os
    from epi_mcmc_b200.model import load_dataset
    from epi_mcmc_b200.sampler import Sampler
    from oracle import cpu_mcmc
    ds = load_dataset("A300")
    M = _model("A300", 4)
    cpr = 4096
    S = Sampler(M, 2 * cpr, kind="demc", seed=1, rungs=2)       # rung 0: beta = 0, rung 1: beta = 1
    S.adapt(True)
    for _ in range(300):
        S.run(10); S.refresh_population()
    S.adapt(False)
    S.record(50, 20)
    S.run(1000)
    dg = S.draws()[0][:, :cpr, :].reshape(-1, M.d)
    dc, info = cpu_mcmc.run(oracle, ds, 4, 256, 2500, 1500, 50, 3, n_threads=os.cpu_count() or 4, beta=0.0, budget_s=120.0)
    dc = dc.reshape(-1, M.d)
    mg, sg = _robust(dg)
    mc, sc = _robust(dc)
    z = np.abs(mg - mc) / sc
    ratio = sg / sc
    print("cpu sampler", info, "z", np.round(z, 2), "spread ratio", np.round(ratio, 2))
    assert info["accept_rate"] > 0.05
    assert (z < 0.5).all() and np.median(z) < 0.15, np.round(z, 2)   # (Monte-Carlo error of the two runs: max z 0.35 observed)
    assert (np.abs(ratio - 1) < 0.3).all(), np.round(ratio, 2)
    # the cold rung is still the posterior's neighbourhood: far narrower than the prior in the identified parameters
    cold = S.draws()[0][:, cpr:, :].reshape(-1, M.d)
    assert (cold.std(axis=0)[:9] < 0.5 * dg.std(axis=0)[:9]).all()
    # the drjacoby-style kernel on the same rung of the same model: one-parameter moves on the reparameterised box
    # (45 Jacobian terms, 45 bandwidths per chain) must land on the same 45 marginals
    D = Sampler(M, 2 * 2048, kind="drj", seed=2, rungs=2)
    D.adapt(True); D.run(120); D.adapt(False)
    D.record(4, 20); D.run(80)
    dd = D.draws()[0][:, :2048, :].reshape(-1, M.d)
    md, sd_ = _robust(dd)
    zd = np.abs(md - mc) / sc
    print("drj prior rung: z", np.round(zd, 2), "spread ratio", np.round(sd_ / sc, 2))
    assert (zd < 0.6).all() and np.median(zd) < 0.2, np.round(zd, 2)   # (2,048 chains x 20 draws: max 0.50 observed)
    assert (np.abs(sd_ / sc - 1) < 0.3).all(), np.round(sd_ / sc, 2)
    M.close()


def test_shipped_series_statistical_parity(oracle):
    """The same check on data the reference ships: the German ECDC series (R/data/ecdc/ecdc.csv through
    R/data/ecdc/data.R:92-98, committed as datasets/DE.json), simple model: GPU DE-MC against the CPU Metropolis
    sampler on the oracle density, and the drjacoby-style kernel keeps that posterior when started from it."""
    from epi_mcmc_b200.model import load_dataset
    from epi_mcmc_b200.sampler import Sampler, run_mcmc
    ds = load_dataset("DE")
    M = _model("DE", 2)
    r1 = run_mcmc(M, burnin=8000, samples=4000, chains=2048, kind="demc", seed=1, thin=40)["draws"]
    d1 = r1.reshape(-1, M.d)
    d3 = _cpu_reference_sampler(oracle, ds, 2, n_chains=256, burnin=4000, samples=3000, thin=10, seed=3)
    m1, s1 = _robust(d1)
    m3, s3 = _robust(d3)
    z = np.abs(m1 - m3) / s3
    print("DE: z", np.round(z, 2), "spread ratio", np.round(s1 / s3, 2), "sd ratio", np.round(d1.std(axis=0) / d3.std(axis=0), 2))
    assert (z < 0.3).all(), z
    assert (np.abs(s1 / s3 - 1) < 0.35).all(), s1 / s3
    T = Sampler(M, 2048, kind="drj", seed=2, theta0=r1[-1])
    T.adapt(True); T.run(150); T.adapt(False)
    T.record(10, 30); T.run(300)
    d2 = T.draws()[0].reshape(-1, M.d)
    z = np.abs(d2.mean(axis=0) - d1.mean(axis=0)) / d1.std(axis=0)
    assert (z < 0.15).all(), z
    assert (np.abs(d2.std(axis=0) / d1.std(axis=0) - 1) < 0.15).all()
    M.close()
