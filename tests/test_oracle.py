"""CPU tests of the oracle (no GPU): the nmath restatements against scipy/mpmath, the RHS
against the reference's own compiled model-agen.cpp, the whole log-posterior against an
independent numpy/scipy transliteration of the R model files (same fixed-step scheme: index
semantics to ~1e-10; scipy LSODA at deSolve's defaults: solver tolerance), and the committed
golden vectors."""
import json
import os

import numpy as np
import pytest

import r_semantics as R
from conftest import DATASETS, ROOT, load_golden, nan_array, rel_err


def ds_load(name):
    with open(os.path.join(ROOT, "epi_mcmc_b200", "datasets", name + ".json")) as f:
        return json.load(f)


def test_pgamma_against_mpmath_and_scipy(oracle):
    import mpmath as mp
    from scipy import special
    mp.mp.dps = 40
    L = oracle.lib()
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(200):
        mean, sd = rng.uniform(4, 25), rng.uniform(2, 9)
        scale = sd * sd / mean
        shape = mean / scale
        q = rng.uniform(0.5, mean + 3 * sd + 1)
        exact = float(mp.gammainc(shape, 0, q / scale, regularized=True))
        worst = max(worst, abs(L.oracle_pgamma(q, shape, scale) - exact) / exact)
        assert abs(special.gammainc(shape, q / scale) - exact) / exact < 1e-12
    assert worst < 5e-13
    assert L.oracle_pgamma(-0.5, 4.0, 2.0) == 0.0  # F(k - 0.5) with k = 0, model-age-groups2q.R:31


def test_densities_against_scipy(oracle):
    from scipy import stats
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(200):
        x = float(rng.integers(0, 3000))
        size = float(rng.choice([0.01, 1, 20, 40, 120, 240]))
        mu = rng.uniform(0.001, 3000)
        ref = stats.nbinom.logpmf(x, size, size / (size + mu))
        assert abs(L.oracle_dnbinom_mu_log(x, size, mu) - ref) <= 1e-9 * max(1.0, abs(ref))
        v, m, s = rng.uniform(0.01, 50), rng.uniform(-3, 30), rng.uniform(0.05, 10)
        assert abs(L.oracle_dnorm_log(v, m, s) - stats.norm.logpdf(v, m, s)) < 1e-12 * max(1, abs(stats.norm.logpdf(v, m, s)))
        assert abs(L.oracle_dlnorm_log(v, m / 10, s / 10) - stats.lognorm.logpdf(v, s=s / 10, scale=np.exp(m / 10))) < 1e-10
        assert abs(L.oracle_pnorm(v, m, s) - stats.norm.cdf(v, m, s)) < 1e-15
    assert L.oracle_dnbinom_mu_log(0.0, 5.0, 2.0) == pytest.approx(5 * np.log(5 / 7), rel=1e-14)
    assert L.oracle_dnbinom_mu_log(1.5, 5.0, 2.0) == -np.inf    # non-integer x: R returns -Inf
    assert np.isnan(L.oracle_dnbinom_mu_log(3.0, 5.0, np.nan))  # NA propagates


def test_gamma_profile_semantics(oracle):
    """calcGammaProfile (model-age-groups2q.R:17-40): extents, normalisation, rev() (an
    asymmetric kernel shows the mirror), against the numpy transliteration."""
    for mean, sd in [(10, 5), (21, 5), (4, 5), (15, 9), (10.3, 2.2), (6, 9), (25, 2)]:
        v, kb, ke = oracle.profile(0, mean, sd)
        ref = R.Age2q.calcGammaProfile(mean, sd)
        assert kb == ref["kbegin"] and ke == ref["kend"] and len(v) == len(ref["values"])
        assert abs(v.sum() - 1) < 1e-14
        assert np.abs(v - ref["values"]).max() < 1e-13
    v, kb, ke = oracle.profile(0, 4.0, 5.0)   # shape 0.64: mass piles up at delay 0..1
    assert kb == -19 and ke == 0 and v[-1] > v[0]   # reversed: the LAST tap holds P(kbegin = 0)
    with pytest.raises(ValueError):
        oracle.profile(0, 10.0, 40.0)             # 241 taps > EPI_MAX_TAPS


def test_normal_profile_semantics(oracle):
    """calcConvolveProfile (model-simple-ode-drj-cmp.R:15-34), called with the negated latency"""
    for lat in (5.0, 10.0, 20.2, 30.0, 40.0):
        v, kb, ke = oracle.profile(1, -lat, 5.0)
        ref = R.Simple.calcConvolveProfile(-lat, 5.0)
        assert kb == ref["kbegin"] and ke == ref["kend"]
        assert np.abs(v - ref["values"]).max() < 1e-14


def test_rhs_matches_reference_binary(oracle):
    """The restated derivs() against the reference's own model-agen.cpp compiled from where it
    lies (oracle/_ref, built by oracle/Makefile): bit-exact, including non-monotone breakpoints
    (index-priority if-chain), t exactly on a breakpoint and the bounds guard."""
    try:
        ref = oracle.RefRHS()
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(3)
    gamma = oracle.lib().oracle_age_gamma()
    tidx = [5, 6, 7, 17, 21, 25, 29, 33, 37, 41]
    for trial in range(600):
        parms = np.zeros(45)
        parms[:5] = [8.55e6, 2.95e6, 1 / 3, 1, gamma]
        ts = np.sort(rng.uniform(30, 330, 10)) if trial % 3 else rng.uniform(30, 330, 10)
        for k, i in enumerate(tidx):
            parms[i] = ts[k]
        for i in range(8, 45):
            if i not in tidx:
                parms[i] = rng.uniform(0.01, 4)
        y = np.concatenate([rng.dirichlet(np.ones(5)) * 8.55e6, rng.dirichlet(np.ones(5)) * 2.95e6])
        if trial % 50 == 0:
            y[rng.integers(10)] = -1.0
        t = rng.uniform(0, 360) if trial % 7 else ts[rng.integers(10)]
        ref.init(parms)
        a, yo = ref.derivs(t, y)
        assert np.array_equal(a, oracle.derivs_age(parms, t, y))
        if y.min() >= 0:  # (when the guard trips the reference leaves yout stale)
            assert np.array_equal(yo[:4], oracle.yout_age(parms, t, y), equal_nan=True)


def test_rhs_agef_matches_reference_binary(oracle):
    """Same for the 8-breakpoint RHS of the secondary age model (model-agef.cpp, parms[37])."""
    try:
        ref = oracle.RefRHS("model-agef", 37)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(4)
    gamma = oracle.lib().oracle_age_gamma()
    tidx = [5, 6, 7, 17, 21, 25, 29, 33]
    for trial in range(600):
        parms = np.zeros(37)
        parms[:5] = [8.55e6, 2.95e6, 1 / 3, 1, gamma]
        ts = np.sort(rng.uniform(30, 330, 8)) if trial % 3 else rng.uniform(30, 330, 8)
        for k, i in enumerate(tidx):
            parms[i] = ts[k]
        for i in range(8, 37):
            if i not in tidx:
                parms[i] = rng.uniform(0.01, 4)
        y = np.concatenate([rng.dirichlet(np.ones(5)) * 8.55e6, rng.dirichlet(np.ones(5)) * 2.95e6])
        if trial % 50 == 0:
            y[rng.integers(10)] = -1.0
        t = rng.uniform(0, 360) if trial % 7 else ts[rng.integers(8)]
        ref.init(parms)
        a, yo = ref.derivs(t, y)
        assert np.array_equal(a, oracle.derivs_agef(parms, t, y))
        if y.min() >= 0:
            assert np.array_equal(yo[:4], oracle.yout_age(parms, t, y), equal_nan=True)


def test_rhs_ager_matches_reference_binary(oracle):
    """The RHS of the later model generation (model-ager.cpp, parms[60]: waning immunity, effective-susceptible cap,
    13-breakpoint schedule, uncertain()) and its nine output variables, incl. the infection incidences y.i / o.i that
    model-age-groups2x.R convolves: the restatement is bit-exact against the reference binary."""
    try:
        ref = oracle.RefRHS("model-ager", 60)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(5)
    gamma = oracle.lib().oracle_age_gamma()
    tidx = [8, 9, 10] + [20 + 4 * k for k in range(10)]
    for trial in range(800):
        parms = np.zeros(60)
        parms[:8] = [8.55e6, 2.95e6, 1 / 3, 1, gamma, 1 / (0.75 * 356), rng.uniform(100, 400), 1.0 if trial % 4 else rng.uniform(0.7, 1.4)]
        ts = np.sort(rng.uniform(30, 330, 13)) if trial % 3 else rng.uniform(30, 330, 13)
        for k, i in enumerate(tidx):
            parms[i] = ts[k]
        for i in range(11, 60):
            if i not in tidx:
                parms[i] = rng.uniform(0.01, 4)
        y = np.concatenate([rng.dirichlet(np.ones(5)) * 8.55e6, rng.dirichlet(np.ones(5)) * 2.95e6])
        if trial % 5 == 0:   # younger susceptibles below the 20 % floor of the effective-susceptible cap
            y[0] = rng.uniform(0, 0.25) * 8.55e6 * 0.5
        if trial % 50 == 0:
            y[rng.integers(10)] = -1.0
        t = rng.uniform(0, 360) if trial % 7 else ts[rng.integers(13)]
        ref.init(parms)
        a, yo = ref.derivs(t, y, nout=9)
        b, yb = oracle.derivs_ager(parms, t, y)
        assert np.array_equal(a, b), (trial, a, b)
        if y.min() >= 0 and (y[:5] <= 8.55e6).all() and (y[5:] <= 2.95e6).all():
            assert np.array_equal(yo[:9], yb, equal_nan=True), (trial, yo[:9], yb)


@pytest.mark.parametrize("name", ["A300", "I290", "X300", "S120", "NE120", "DE"])
def test_oracle_vs_numpy_transliteration_same_scheme(name, oracle):
    """Index semantics (filter alignment + mirror, two-pass offset, rate-map slices, dnbinom
    recycling, window clamps, NA and invalid-offset branches) against an independent
    vector-style transliteration of the R files using the same fixed-step scheme."""
    ds = ds_load(name)
    g = load_golden(name)
    theta = np.array(g["theta"])
    idx = list(range(0, len(theta), 3))[:14]
    res = oracle.evaluate(ds["flavour"], ds, theta[idx], ds["period"], 4)
    for k, i in enumerate(idx):
        if ds["flavour"] in ("age2q", "age2i", "age2x"):
            M = {"age2q": R.Age2q, "age2i": R.Age2i, "age2x": R.Age2x}[ds["flavour"]](ds, ds["period"], "rk4", 4)
            try:
                ll, terms, st = M.calclogl(theta[i])
            except AssertionError:   # (age-2x: a symptomatic-testing window beyond the simulated period)
                assert res["status"][k] & 16
                continue
        else:
            M = R.Simple(ds, ds["period"], "rk4", 4, ne=ds["flavour"] == "ne")
            p = theta[i].copy()
            p[4] = np.exp(p[4])
            ll, terms, st = M.calclogl(p)
        lp = M.calclogp(theta[i])
        assert st["padding"] == res["padding"][k]
        assert (res["offset"][k] == 10000) == bool(res["status"][k] & 1)
        if not res["status"][k] & 1:
            assert st["offset"] == res["offset"][k]
        assert rel_err(ll, res["logl"][k]) < 2e-9     # scipy's nbinom/gammainc carry ~1e-10 of their own
        assert rel_err(lp, res["logp"][k]) < 1e-13


@pytest.mark.parametrize("name,tol_m4,tol_lsoda", [("A300", 0.02, 0.05), ("I290", 0.02, 0.05), ("S120", 1e-3, 1e-2)])
def test_fixed_step_scheme_vs_lsoda(name, tol_m4, tol_lsoda, oracle):
    """The oracle's breakpoint-cut RK4 (m = 4) against scipy LSODA: (a) at deSolve's defaults
    rtol = atol = 1e-6, hmax = 1 (the reference's own solver noise) and (b) tight.  At the model
    file's init the m = 4 scheme is as close to the tight solution as LSODA(1e-6) is."""
    ds = ds_load(name)
    g = load_golden(name)
    th = np.array(g["theta"])[0]
    o4 = oracle.evaluate(ds["flavour"], ds, th[None], ds["period"], 4)["logl"][0]
    if ds["flavour"] in ("age2q", "age2i"):
        cls = R.Age2q if ds["flavour"] == "age2q" else R.Age2i
        mk = lambda integ: cls(ds, ds["period"], integ, 4).calclogl(th)[0]
    else:
        p = th.copy()
        p[4] = np.exp(p[4])
        mk = lambda integ: R.Simple(ds, ds["period"], integ, 4).calclogl(p)[0]
    tight, loose = mk("lsoda_tight"), mk("lsoda")
    assert abs(o4 - tight) < tol_m4
    assert abs(loose - tight) < tol_lsoda
    assert abs(o4 - tight) <= max(abs(loose - tight), 1e-4) * 1.5


@pytest.mark.parametrize("name", DATASETS)
def test_golden_vectors_reproduce(name, oracle):
    """The committed goldens are what the oracle computes today (guards against silent drift)."""
    ds = ds_load(name)
    g = load_golden(name)
    theta = np.array(g["theta"])
    for m in ("1", "4"):
        r = oracle.evaluate(ds["flavour"], ds, theta, ds["period"], int(m), n_threads=4)
        res = g["results"][m]
        assert np.array_equal(r["status"], res["status"]) and np.array_equal(r["offset"], res["offset"])
        assert np.array_equal(r["padding"], res["padding"])
        assert rel_err(r["logl"], nan_array(res["logl"])).max() < 1e-13
        assert rel_err(r["logp"], np.array(res["logp"])).max() < 1e-14


def test_oracle_threads_agree(oracle):
    ds = ds_load("S120")
    th = np.array(load_golden("S120")["theta"])
    a = oracle.evaluate("simple", ds, th, 120, 4, n_threads=1)
    b = oracle.evaluate("simple", ds, th, 120, 4, n_threads=8)
    assert np.array_equal(a["logl"], b["logl"], equal_nan=True) and np.array_equal(a["status"], b["status"])


def test_invalid_offset_and_na_branches(oracle):
    ds = ds_load("A300")
    mn, mx, init = oracle.df_params("age2q", ds, 300)
    th = np.tile(init, (3, 1))
    th[0, 0:3] = mn[0:3]
    th[0, 17] = mx[17]
    th[1, [11, 15]] = 15.0
    th[1, 19] = 9.0
    th[1, [12, 16]] = 10.0
    th[1, 20] = 2.0
    th[1, [43, 44]] = 4.0
    th[1, 17] = 0.0
    th[2, 19] = 40.0
    r = oracle.evaluate("age2q", ds, th, 300, 4)
    assert r["offset"][0] == 10000 and r["status"][0] & 1 and np.isfinite(r["logl"][0])
    assert r["status"][1] & 2 and np.isnan(r["logl"][1])
    assert r["status"][2] & 16 and np.isnan(r["logl"][2])


def test_ode_path_through_the_reference_binary(oracle):
    """The whole pass-2 ODE path pinned to the reference's compiled code: the parms vector built the way
    model-age-groups2q.R:193-207 builds it (position-coupled to the #defines of model-agen.cpp:7-51) is handed to
    the reference's own derivs() (oracle/_ref/model-agen.so) and integrated with scipy LSODA at tight tolerances
    from the R file's initial state; the oracle's calculateModel states (breakpoint-cut RK4, m = 8 and 16) must converge
    to it at fourth order on every compartment and every day.  A wrong slot in parms, a wrong segment
    kind or a wrong initial state shows up as an O(1) difference."""
    from scipy import integrate
    try:
        ref = oracle.RefRHS()
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ds = ds_load("A300")
    P = ds["period"]
    for th in np.array(load_golden("A300")["theta"])[[0, 3, 5]]:
        series, states, off, pad = oracle.trajectory("age2q", ds, th, P, 8)
        states16 = oracle.trajectory("age2q", ds, th, P, 16)[1]
        if off == 10000:
            continue
        p = lambda i: th[i - 1]   # 1-based like the R file (:52-99)
        lo, ltp = ds["lockdown_offset"], ds["lockdown_transition_period"]
        t2 = off + lo + ltp; t3 = t2 + p(24); t4 = t3 + p(25); t5 = off + ds["d5"]; t6 = t5 + p(32); t7 = t6 + p(36)
        t8 = off + ds["d8"]; t9 = t8 + 7
        gamma = oracle.lib().oracle_age_gamma()
        b = {0: (p(1), p(2), p(3)), 1: (p(4), p(5), p(6)), 2: (p(7), p(8), p(9)), 3: (p(7), p(27), p(28)),
             4: (p(26), p(27), p(28)), 5: (p(29), p(30), p(31)), 6: (p(33), p(34), p(35)), 7: (p(37), p(38), p(39)),
             8: (p(37), p(38), p(39)), 9: (p(40) * p(37), p(41) * p(38), p(42) * p(39))}
        parms = [ds["y_N"], ds["o_N"], 1 / 3, 1.0, gamma, off + lo + p(19), off + lo + max(p(19), 0), t2,
                 *b[0], *b[1], *b[2]]
        for k, tk in zip(range(3, 10), (t3, t4, t5, t6, t7, t8, t9)):
            parms += [tk, *b[k]]
        ref.init(np.array(parms))
        Y0 = [ds["y_N"] - 1, 1, 0, 0, 0, ds["o_N"], 0, 0, 0, 0]
        times = np.arange(pad + 1, pad + P + 1, dtype=float)
        sol = integrate.solve_ivp(lambda t, y: ref.derivs(t, y)[0], (times[0], times[-1]), Y0, method="LSODA",
                                  t_eval=times, rtol=1e-12, atol=1e-9, max_step=0.1)
        assert sol.success and sol.y.shape[1] == P
        got = sol.y.T
        # states: [day][Sy, E1y, E2y, Iy, Ry, So, E1o, E2o, Io, Ro]; persons below 1 compared absolutely
        scale = np.maximum(np.abs(got), 1.0)
        e8, e16 = ((np.abs(st - got) / scale).max() for st in (states, states16))
        # fourth-order convergence TOWARDS the reference binary's solution: 16x per halving of the step
        assert e8 < 1e-4 and e16 < 1e-5 and e16 < e8 / 8, (e8, e16)


def test_ode_path_2i_through_the_reference_binary(oracle):
    """Same for the secondary age model: parms as model-age-groups2i.R:176-191 builds them from params[1..36]
    (:52-93), the reference's model-agef.so, fourth-order convergence of the oracle's states towards it."""
    from scipy import integrate
    try:
        ref = oracle.RefRHS("model-agef", 37)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ds = ds_load("I290")
    P = ds["period"]
    checked = 0
    for th in np.array(load_golden("I290")["theta"])[:6]:
        series, states, off, pad = oracle.trajectory("age2i", ds, th, P, 8)
        if off == 10000:
            continue
        states16 = oracle.trajectory("age2i", ds, th, P, 16)[1]
        p = lambda i: th[i - 1]
        lo, ltp = ds["lockdown_offset"], ds["lockdown_transition_period"]
        t2 = off + lo + ltp; t3 = t2 + p(25); t4 = t3 + p(26); t5 = off + ds["d5"]; t6 = t5 + p(30); t7 = off + ds["d7"]
        b2, b4, b6 = (p(7), p(8), p(9)), (p(27), p(28), p(29)), (p(31), p(32), p(33))
        b7 = (p(34) * p(31), p(36) * p(32), p(35) * p(33))     # (betay7, betao7, betayo7): note fo7 = params[36]
        parms = [ds["y_N"], ds["o_N"], 1 / 3, 1.0, oracle.lib().oracle_age_gamma(), off + lo + p(21),
                 off + lo + max(p(21), 0), t2, p(1), p(2), p(3), p(4), p(5), p(6), *b2,
                 t3, *b2, t4, *b4, t5, *b4, t6, *b6, t7, *b7]
        ref.init(np.array(parms))
        Y0 = [ds["y_N"] - 1, 1, 0, 0, 0, ds["o_N"], 0, 0, 0, 0]
        times = np.arange(pad + 1, pad + P + 1, dtype=float)
        sol = integrate.solve_ivp(lambda t, y: ref.derivs(t, y)[0], (times[0], times[-1]), Y0, method="LSODA",
                                  t_eval=times, rtol=1e-12, atol=1e-9, max_step=0.1)
        assert sol.success and sol.y.shape[1] == P
        got = sol.y.T
        scale = np.maximum(np.abs(got), 1.0)
        e8, e16 = ((np.abs(st - got) / scale).max() for st in (states, states16))
        assert e8 < 1e-4 and e16 < 1e-5 and e16 < e8 / 8, (e8, e16)
        checked += 1
    assert checked >= 2


def test_ode_path_2x_through_the_reference_binary(oracle):
    """The later generation: parms as model-age-groups2x.R:222-238 builds them (60 values, position-coupled to
    model-ager.cpp:7-66), the reference's model-ager.so under scipy LSODA, fourth-order convergence of the oracle's
    states AND of its incidence series y.i / o.i (output variables 5 and 6, what the observation models convolve)."""
    from scipy import integrate
    try:
        ref = oracle.RefRHS("model-ager", 60)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ds = ds_load("X300")
    P = ds["period"]
    checked = 0
    for th in np.array(load_golden("X300")["theta"])[[0, 26, 27, 30]]:
        series, states, off, pad = oracle.trajectory("age2x", ds, th, P, 8)
        if off == 10000:
            continue
        states16 = oracle.trajectory("age2x", ds, th, P, 16)[1]
        p = lambda i: th[i - 1]   # 1-based like the R file (:56-119)
        lo, ltp = ds["lockdown_offset"], ds["lockdown_transition_period"]
        t2 = off + lo + ltp; t3 = t2 + p(24); t4 = t3 + p(28); t5 = off + ds["d5"]; t6 = t5 + p(35); t7 = t6 + p(39)
        t8 = off + ds["d8"]; t9 = off + ds["d9"]; t10 = off + p(49); t11 = t10 + 3; t12 = off + ds["d12"]
        b = {0: (p(1), p(2), p(3)), 1: (p(4), p(5), p(6)), 2: (p(7), p(8), p(9)), 3: (p(25), p(26), p(27)),
             4: (p(29), p(30), p(31)), 5: (p(32), p(33), p(34)), 6: (p(36), p(37), p(38)), 7: (p(40), p(41), p(42)),
             8: (p(40), p(41), p(42)), 9: (p(43), p(44), p(45)), 10: (p(43), p(44), p(45)), 11: (p(50), p(51), p(52)),
             12: (p(43), p(44), p(45))}
        gamma = oracle.lib().oracle_age_gamma()
        parms = [ds["y_N"], ds["o_N"], 1 / 3, 1.0, gamma, 1 / (0.75 * 356), off + ds["duncertain"], 1.0,
                 off + lo + p(19), off + lo + max(p(19), 0), t2, *b[0], *b[1], *b[2]]
        for k, tk in zip(range(3, 13), (t3, t4, t5, t6, t7, t8, t9, t10, t11, t12)):
            parms += [tk, *b[k]]
        assert len(parms) == 60
        ref.init(np.array(parms))
        Y0 = [ds["y_N"] - 1, 1, 0, 0, 0, ds["o_N"], 0, 0, 0, 0]
        times = np.arange(pad + 1, pad + P + 1, dtype=float)
        sol = integrate.solve_ivp(lambda t, y: ref.derivs(t, y, nout=9)[0], (times[0], times[-1]), Y0, method="LSODA",
                                  t_eval=times, rtol=1e-12, atol=1e-9, max_step=0.1)
        assert sol.success and sol.y.shape[1] == P
        got = sol.y.T
        scale = np.maximum(np.abs(got), 1.0)
        e8, e16 = ((np.abs(st - got) / scale).max() for st in (states, states16))
        assert e8 < 1e-4 and e16 < 1e-5 and e16 < e8 / 8, (e8, e16)
        checked += 1
    assert checked >= 2


def test_division_by_six_through_fma_is_the_division():
    """The ODE kernels compute hh / 6 of a cut RK4 piece as q = x * RN(1/6); r = fma(-6, q, x); q + r * RN(1/6)
    (epi_kernels.cu:div6) while the oracle divides: the two must agree in every bit (proof in the kernel source;
    here 10^6 random piece lengths, values next to powers of two and next to multiples of six ulps)."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    z = 1.0 / 6.0
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(0.0, 0.25, 400000), 10.0 ** rng.uniform(-17, 2, 400000),
                         np.ldexp(rng.integers(1, 1 << 53, 200000).astype(np.float64), -rng.integers(50, 70, 200000)),
                         np.nextafter(2.0 ** np.arange(-40.0, 4.0), 0.0), np.nextafter(2.0 ** np.arange(-40.0, 4.0), 9.0),
                         2.0 ** np.arange(-40.0, 4.0), [0.0, 0.25, 0.125, 1.0, 6.0, 3.0]])
    bad = 0
    for x in xs.tolist():
        q = x * z
        r = libm.fma(-6.0, q, x)
        bad += libm.fma(r, z, q) != x / 6.0
    assert bad == 0
