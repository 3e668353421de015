"""GPU parity: the CUDA path through the C ABI (epi_eval / epi_eval_detail /
epi_trajectories) against (a) the committed golden vectors and (b) the CPU
oracle run live on seeded inputs.  Bar (BASELINE.json north_star): logl + logp
within 1e-10 RELATIVE in FP64 at fixed theta; integer outputs (data_offset,
padding, status) bit-exact.

Ill-conditioned theta: the reference algorithm forms incidence as differences of
N - convolution(S); once an epidemic has burnt out (far-out box corners, logl ~ -1e5)
those differences cancel catastrophically and ANY two correctly rounded FP64
implementations differ by more than 1e-10.  The bar used here is therefore
max(1e-10 * |logl|, 4 x the oracle's own change when its ODE states are jittered by
<= 32 ulps, oracle.conditioning) — for every theta near the posterior mass the first term is the binding one,
which test_well_conditioned_strict asserts without the second term."""
import numpy as np
import pytest

from conftest import DATASETS, load_golden, logl_within_bar, nan_array, oracle_conditioning, rel_err

pytestmark = pytest.mark.gpu

TOL_LOGL = 1e-10   # the north_star tolerance, relative
TOL_LOGP = 1e-12


def _model(name, m):
    from epi_mcmc_b200.model import Model, load_dataset
    return Model(load_dataset(name), substeps=m)


@pytest.mark.parametrize("name", DATASETS)
@pytest.mark.parametrize("m", [1, 4])
def test_golden_vectors(name, m):
    g = load_golden(name)
    theta = np.array(g["theta"])
    res = g["results"][str(m)]
    M = _model(name, m)
    logl, logp, status = M.calclogpost_batch(theta)
    off, pad, terms = M.eval_detail(theta)
    assert np.array_equal(status, np.array(res["status"], dtype=np.int32))
    assert np.array_equal(off, np.array(res["offset"], dtype=np.int32))
    assert np.array_equal(pad, np.array(res["padding"], dtype=np.int32))
    ref_l, ref_p = nan_array(res["logl"]), np.array(res["logp"])
    cond = np.array(res["cond"])
    within = logl_within_bar(logl, ref_l, cond, TOL_LOGL)
    assert within.all(), [(i, logl[i], ref_l[i], cond[i]) for i in np.nonzero(~within)[0]]
    assert rel_err(logp, ref_p).max() <= TOL_LOGP
    assert logl_within_bar(logl + logp, ref_l + ref_p, cond, TOL_LOGL).all()
    # theta[0] is the model file's init: the strict bar, no conditioning allowance
    assert rel_err(logl[:1] + logp[:1], ref_l[:1] + ref_p[:1]).max() <= TOL_LOGL
    nt = 6 if g["flavour"] in ("age2q", "age2i") else 3
    ref_t = np.array([[np.nan if v is None else v for v in row] for row in res["terms"]])[:, :nt]
    # individual terms can be tiny next to logl: compare them on the scale of the total
    scale = np.where(np.isfinite(ref_l), np.abs(ref_l), np.nanmax(np.abs(np.where(np.isfinite(ref_t), ref_t, 0.0)), axis=1))
    bar = np.maximum(TOL_LOGL * np.maximum(scale, 1.0), 4.0 * cond)[:, None]
    with np.errstate(invalid="ignore"):
        d = np.abs(terms[:, :nt] - ref_t) / bar
    d = np.where(np.isnan(terms[:, :nt]) & np.isnan(ref_t), 0.0, d)
    d = np.where((terms[:, :nt] == ref_t), 0.0, d)
    assert np.nanmax(d) <= 1.0 and not np.isnan(d).any()
    M.close()


@pytest.mark.parametrize("name,n", [("S120", 2048), ("NE120", 2048), ("A300", 1024), ("I290", 1024), ("X300", 1024), ("SE_NE", 1024)])
def test_oracle_live_random_theta(name, n, oracle):
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(7)
    box = mn + rng.uniform(size=(n // 2, M.d)) * (mx - mn)
    near = np.clip(init + (rng.uniform(size=(n - n // 2, M.d)) - 0.5) * 0.3 * (mx - mn), mn, mx)
    theta = np.vstack([box, near])
    logl, logp, status = M.calclogpost_batch(theta)
    off, pad, _ = M.eval_detail(theta)
    ref = oracle.evaluate(ds["flavour"], ds, theta, ds["period"], 4, n_threads=8)
    # a theta whose died(t) touches the threshold within rounding may legitimately pick the
    # neighbouring day, and one whose hospital series are pure cancellation residue in the ratio window
    # (no epidemic: ytotal/(ytotal + ototal) is noise/noise or 0/0, :563-565) may flip NA <-> finite — a wide
    # sweep (tests/_sweep.py, 10^5 box-uniform draws) sees the latter at ~1e-4, only in that term:
    # allow at most 1 in 1000, everything else must match exactly
    mism = (off != ref["offset"]) | (status != ref["status"]) | (pad != ref["padding"])
    assert mism.sum() <= max(1, n // 1000), mism.sum()
    ok = ~mism
    cond = oracle_conditioning(oracle, ds["flavour"], ds, theta, ds["period"], 4, ref["logl"])
    within = logl_within_bar(logl[ok], ref["logl"][ok], cond[ok], TOL_LOGL)
    # Box-uniform theta reach regimes where the reference arithmetic itself is noise (burnt-out
    # epidemics: the hospital y/o ratio term divides two sums of cancellation residue; explosive
    # beta*h > 2.8 where RK4 amplifies rounding): at most 1 % may miss the bar, none by much.
    assert within.mean() >= 0.99, (int((~within).sum()), n)
    assert (rel_err(logl[ok], ref["logl"][ok]) <= TOL_LOGL).mean() > 0.85
    assert rel_err(logl[ok], ref["logl"][ok]).max() < 1e-3
    assert rel_err(logp[ok], ref["logp"][ok]).max() <= TOL_LOGP
    assert (status[ok] == 0).sum() > n // 4   # the comparison is not vacuous
    M.close()


@pytest.mark.parametrize("name,n", [("S365", 512), ("A300", 512), ("A365", 256), ("I290", 512), ("X300", 512)])
def test_well_conditioned_strict(name, n, oracle):
    """theta close to the model file's init (where the posterior
    mass of the synthetic data sits; +-0.5 % of the box width: the north_star bar holds as stated, 1e-10 relative on
    logl + logp, with no conditioning allowance."""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(11)
    theta = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.01 * (mx - mn), mn, mx)
    logl, logp, status = M.calclogpost_batch(theta)
    ref = oracle.evaluate(ds["flavour"], ds, theta, ds["period"], 4, n_threads=8)
    assert np.array_equal(status, ref["status"])
    good = status == 0
    assert good.sum() > n // 2
    assert rel_err((logl + logp)[good], (ref["logl"] + ref["logp"])[good]).max() <= TOL_LOGL
    M.close()


@pytest.mark.parametrize("name", ["A300", "I290", "X300", "S365", "NE120"])
def test_mirrored_oracle_states_are_bit_identical(name, oracle):
    """oracle_set_mirror(1): the oracle evaluates the fixed-step scheme in the kernels' operation order (curves
    pre-divided by the population, explicit fma everywhere).  Every operation is correctly rounded IEEE on both
    sides, so the ODE state series the GPU reports (S, and E / I / R of the ext trajectories) must then equal the
    oracle's BIT FOR BIT — over box-wide theta, breakpoints inside sub-steps, guard hits and all."""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(23)
    theta = np.vstack([init[None], np.clip(init + (rng.uniform(size=(24, M.d)) - 0.5) * 0.3 * (mx - mn), mn, mx),
                       mn + rng.uniform(size=(24, M.d)) * (mx - mn)])
    model_th = M.transformParams(theta)
    series, off, pad = M.calculateModel_batch(model_th, ds["period"], ext=True)
    names = M.series_names_ext
    state_cols = [k for k, nm in enumerate(names) if nm.split(".")[-1] in ("S", "E", "I", "R")]
    assert len(state_cols) >= 4
    oracle.set_mirror(True)
    try:
        n_checked = 0
        for i in range(len(theta)):
            try:
                ref, _, roff, rpad = oracle.trajectory(ds["flavour"], ds, model_th[i], ds["period"], 4, ext=True)
            except ValueError:
                continue
            if off[i] != roff or pad[i] != rpad:
                continue   # the threshold search reads the death profile too: not what this test is about
            for k in state_cols:
                assert np.array_equal(series[i, k], ref[k], equal_nan=True), (i, names[k], np.abs(series[i, k] - ref[k]).max())
            n_checked += 1
        assert n_checked >= 40
    finally:
        oracle.set_mirror(False)
    M.close()


@pytest.mark.parametrize("name,n", [("A300", 1024), ("I290", 1024), ("S120", 1024)])
def test_box_uniform_against_the_mirrored_oracle(name, n, oracle):
    """With the ODE states bit-identical (mirrored oracle), what is left of a GPU <-> oracle difference in logl is the
    latency profiles (incomplete-gamma series in a different summation order) and the table-driven logarithm: the
    1e-10 bar then holds for (nearly) every box-uniform theta, with no conditioning allowance — the misses of
    test_oracle_live_random_theta against the reference's operation order are the cancellation of
    diff(N - filter(S)) amplifying a-few-ulp differences of the states, not a property of the kernels."""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    mn, mx = M.df_params["min"], M.df_params["max"]
    rng = np.random.default_rng(7)
    theta = mn + rng.uniform(size=(n, M.d)) * (mx - mn)
    logl, logp, status = M.calclogpost_batch(theta)
    off, pad, _ = M.eval_detail(theta)
    oracle.set_mirror(True)
    try:
        ref = oracle.evaluate(ds["flavour"], ds, theta, ds["period"], 4, n_threads=8)
    finally:
        oracle.set_mirror(False)
    ok = (off == ref["offset"]) & (status == ref["status"]) & (pad == ref["padding"])
    assert (~ok).sum() <= max(1, n // 1000)
    fin = ok & np.isfinite(ref["logl"])
    err = rel_err(logl[fin], ref["logl"][fin])
    rate = float((err <= TOL_LOGL).mean())
    print(name, "box-uniform theta within 1e-10 of the mirrored oracle:", rate, "of", int(fin.sum()), "max rel err", err.max())
    # measured: S120 and I290 >= 0.97; A300 0.92 (against 0.85 with the reference's operation order) — its misses are
    # the profile weights, a few ulp apart, amplified by the same cancellation in burnt-out epidemics
    assert rate >= (0.90 if name == "A300" else 0.97), rate
    M.close()


def test_soa_layout_and_ragged_batches():
    """AoS and SoA inputs agree bit for bit; batch sizes that are not multiples of the
    32-chain tile, including n = 1, work; n = 0 is a no-op."""
    M = _model("A300", 4)
    rng = np.random.default_rng(3)
    init, w = M.df_params["init"], M.df_params["max"] - M.df_params["min"]
    theta = np.clip(init + (rng.uniform(size=(77, M.d)) - 0.5) * 0.1 * w, M.df_params["min"], M.df_params["max"])
    l1, p1, s1 = M.calclogpost_batch(theta)
    l2, p2, s2 = M.calclogpost_batch(np.ascontiguousarray(theta.T), layout="soa")
    assert np.array_equal(l1, l2, equal_nan=True) and np.array_equal(p1, p2) and np.array_equal(s1, s2)
    for n in (1, 31, 33):
        l3, p3, _ = M.calclogpost_batch(theta[:n])
        assert np.array_equal(l3, l1[:n], equal_nan=True) and np.array_equal(p3, p1[:n])
    l0, p0, s0 = M.calclogpost_batch(np.zeros((0, M.d)))
    assert l0.size == 0
    # the scalar reference API
    assert M.calclogl(M.transformParams(theta[0])) == l1[0]
    assert M.calclogp(theta[0]) == p1[0]
    M.close()


def test_edge_statuses():
    """InvalidDataOffset (no threshold crossing), NA (hospital kernel wider than the padding),
    out-of-box and unsupported (profile wider than EPI_MAX_TAPS) are flagged like the oracle does."""
    from epi_mcmc_b200 import _capi
    from epi_mcmc_b200.model import load_dataset
    from oracle import oracle
    ds = load_dataset("A300")
    M = _model("A300", 4)
    init = M.df_params["init"].copy()
    th = np.tile(init, (4, 1))
    th[0, 0:3] = M.df_params["min"][0:3]    # slowest epidemic the box allows ...
    th[0, 17] = M.df_params["max"][17]      # ... never reaches the largest t0_morts in 60 days -> invalid offset
    th[1, [11, 15]] = 15.0                  # hosp latency 15, HLsd 9 -> kend 42
    th[1, 19] = 9.0
    th[1, [12, 16]] = 10.0                  # died latency 10, DLsd 2 -> kend 16; case kend <= 25 -> padding 26 < 42
    th[1, 20] = 2.0
    th[1, [43, 44]] = 4.0
    th[1, 17] = 0.0                         # threshold crossed on the first day: the scored window starts inside the NA days
    th[2, 0] = 100.0                        # out of box
    th[3, 19] = 40.0                        # HLsd = 40 -> 241 taps: unsupported
    logl, logp, status = M.calclogpost_batch(th)
    ref = oracle.evaluate("age2q", ds, th, 300, 4)
    assert np.array_equal(status, ref["status"])
    assert status[0] & _capi.ST_INVALID_OFFSET and np.isfinite(logl[0])
    assert status[1] & _capi.ST_NA and np.isnan(logl[1])
    assert status[2] & _capi.ST_OUT_OF_BOX
    assert status[3] & _capi.ST_UNSUPPORTED and np.isnan(logl[3])
    cond = oracle_conditioning(oracle, "age2q", ds, th, 300, 4, ref["logl"])
    assert logl_within_bar(logl, ref["logl"], cond, TOL_LOGL).all()
    M.close()


@pytest.mark.parametrize("name", ["A300", "I290", "X300", "S120", "NE120"])
def test_trajectories_match_oracle(name, oracle):
    """calculateModel(params, period) series, also for a horizon other than FitTotalPeriod
    (quantileData calls it with 3 * period, libMCMC.R:160-161)."""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    g = load_golden(name)
    theta = np.array(g["theta"])[:12]
    model_th = M.transformParams(theta)
    for P in (ds["period"], ds["period"] + 77):
        series, off, pad = M.calculateModel_batch(model_th, P)
        for i in range(len(theta)):
            ref, _, roff, rpad = oracle.trajectory(ds["flavour"], ds, model_th[i], ds["period"], 4, calc_period=P)
            assert off[i] == roff and pad[i] == rpad
            # incidence is a difference of N - filter(S): its absolute noise floor is a few ulp(N)
            Npop = ds["N"]
            bar = 1e-10 * np.nanmax(np.abs(ref), axis=1, keepdims=True) + 16 * 2.0 ** -52 * Npop
            with np.errstate(invalid="ignore"):
                err = np.abs(series[i] - ref) / bar
            err = np.where(np.isnan(series[i]) & np.isnan(ref), 0.0, err)
            assert not np.isnan(err).any() and err.max() <= 1.0, err.max()
    # golden head/tail of the init trajectory
    gi = g["init_trajectory"]
    s0, o0, p0 = M.calculateModel_batch(model_th[:1])
    assert o0[0] == gi["offset"] and p0[0] == gi["padding"]
    floor = 16 * 2.0 ** -52 * ds["N"]
    head = np.array(gi["head"])
    assert (np.abs(s0[0][:, :5] - head) <= 1e-10 * np.abs(head) + floor).all()
    chk = np.array(gi["checksum"])
    assert (np.abs(s0[0].sum(axis=1) - chk) <= 1e-10 * np.abs(chk) + floor * s0.shape[2]).all()
    M.close()


@pytest.mark.parametrize("name,fun", [("A300", ("y.hospi", "o.hospi")), ("A300", "y.deadi"), ("I290", "o.casei"), ("S120", "hospi")])
def test_quantile_data_matches_reference_semantics(name, fun, oracle):
    """quantileData (libMCMC.R:151-172) on the device against the oracle + numpy: per posterior draw
    calculateModel(params, 3 * period), takeAndPad by the draw's data_offset, per-day R type-7
    quantiles with na.rm."""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(5)
    n, period, startoffset = 200, 100, 3
    sample = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.04 * (mx - mn), mn, mx)
    probs = [0.05, 0.5, 0.95]
    got = M.quantileData(sample, fun, startoffset, period, probs)
    names = (fun,) if isinstance(fun, str) else fun
    simperiod = 3 * period
    data = np.full((simperiod, n), np.nan)
    model_th = M.transformParams(sample)
    for i in range(n):
        ser, _, off, pad = oracle.trajectory(ds["flavour"], ds, model_th[i], ds["period"], 4, calc_period=simperiod)
        if off == 10000:
            continue
        vec = np.zeros(pad + simperiod)                      # fun(state): full-length sim-indexed vector
        for nm in names:
            vec[pad:] += ser[M.series_names.index(nm)]
        o = off - startoffset
        take = vec[o - 1:simperiod]                          # takeAndPad(data, offset, l): data[offset:min(l, length)]
        data[:len(take), i] = take
    ref = np.nanquantile(data[:period], probs, axis=1).T     # numpy 'linear' == R type 7
    scale = np.nanmax(np.abs(ref)) + 1e-300
    assert np.abs(got - ref).max() <= 1e-9 * scale + 32 * 2.0 ** -52 * ds["N"]
    M.close()


def test_quantile_data_beyond_the_shared_memory_sort():
    """More than 8,192 posterior draws: the per-day sort runs in a global scratch buffer.  Checked against numpy on the
    GPU's own trajectories (which the tests above compare with the oracle): same takeAndPad alignment, type-7
    quantiles, na.rm."""
    M = _model("S120", 2)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(9)
    n, period, startoffset = 10000, 60, 3
    sample = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.06 * (mx - mn), mn, mx)
    probs = [0.05, 0.25, 0.5, 0.95]
    got = M.quantileData(sample, "hospi", startoffset, period, probs)
    simperiod = 3 * period
    series, off, pad = M.calculateModel_batch(M.transformParams(sample), simperiod)
    data = np.full((simperiod, n), np.nan)
    k = M.series_names.index("hospi")
    for i in range(n):
        if off[i] == 10000 or off[i] - startoffset < 1:
            continue
        vec = np.concatenate([np.zeros(pad[i]), series[i, k]])
        take = vec[off[i] - startoffset - 1:simperiod]
        data[:len(take), i] = take
    ref = np.nanquantile(data[:period], probs, axis=1).T
    assert np.abs(got - ref).max() <= 1e-12 * (np.nanmax(np.abs(ref)) + 1e-300)
    M.close()


@pytest.mark.parametrize("name,fun,margin", [("A300", "Re", ("max", 40, 100)), ("A300", ("y.hospi", "o.hospi"), ("sum", 1, 120)),
                                             ("X300", "y.casei", ("mean", 100, 160)), ("S120", "hospi", ("min", 5, 30))])
def test_marginalize_data_matches_reference_semantics(name, fun, margin, oracle):
    """marginalizeData (libMCMC.R:174-188; predictMCMC.R:714-717 takes the maximum of Re over a window) on the device
    against the oracle + numpy: calculateModel(params, 3 * period), takeAndPad by the draw's data_offset, one number
    per draw, NA when the range holds an NA."""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(6)
    n, period, startoffset = 300, 100, 2
    sample = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.04 * (mx - mn), mn, mx)
    got = M.marginalizeData(sample, fun, startoffset, period, margin)
    names = (fun,) if isinstance(fun, str) else fun
    kind, j1, j2 = margin
    simperiod = 3 * period
    model_th = M.transformParams(sample)
    ref = np.full(n, np.nan)
    for i in range(n):
        ser, _, off, pad = oracle.trajectory(ds["flavour"], ds, model_th[i], ds["period"], 4, calc_period=simperiod, ext=True)
        if off == 10000:
            continue
        vec = np.zeros(pad + simperiod)
        for nm in names:
            vec[pad:] += ser[M.series_names_ext.index(nm)]
        v = np.full(simperiod, np.nan)
        take = vec[off - startoffset - 1:simperiod]
        v[:len(take)] = take
        ref[i] = {"max": np.max, "sum": np.sum, "min": np.min, "mean": np.mean}[kind](v[j1 - 1:j2])
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert ok.sum() > n // 2
    scale = np.abs(ref[ok]).max() + 1e-300
    assert np.abs(got[ok] - ref[ok]).max() <= 1e-9 * scale + 64 * 2.0 ** -52 * ds["N"]
    M.close()


@pytest.mark.parametrize("name,m", [("A300", 4), ("I290", 4), ("S365", 2), ("NE120", 1)])
def test_pass2_restart_is_bit_identical(name, m, monkeypatch):
    """Pass 2 resumes from the pass-1 checkpoint after its first floor(min breakpoint) - t0 days
    (both passes share the fixed-step scheme and beta0 until then): switching the restart off must
    not change a single bit of logl, offsets or trajectories — including theta whose first
    breakpoint lies before the start (t0o = -30) or beyond the checkpointed days."""
    from epi_mcmc_b200.model import Model, load_dataset
    g = load_golden(name)
    theta = np.array(g["theta"])
    ds = load_dataset(name)
    monkeypatch.setenv("EPI_PASS2_RESTART", "1")
    A = Model(ds, substeps=m)
    monkeypatch.setenv("EPI_PASS2_RESTART", "0")
    B = Model(ds, substeps=m)
    la, pa, sa = A.calclogpost_batch(theta)
    lb, pb, sb = B.calclogpost_batch(theta)
    assert np.array_equal(la, lb, equal_nan=True) and np.array_equal(sa, sb)
    ta, oa, _ = A.calculateModel_batch(A.transformParams(theta[:16]), ds["period"] + 40)
    tb, ob, _ = B.calculateModel_batch(B.transformParams(theta[:16]), ds["period"] + 40)
    assert np.array_equal(ta, tb, equal_nan=True) and np.array_equal(oa, ob)
    A.close()
    B.close()


@pytest.mark.parametrize("name,period", [("A300", 301), ("A300", 333), ("I290", 291), ("S120", 121), ("NE120", 127)])
def test_odd_periods(name, period, oracle):
    """FitTotalPeriod values that are odd (rows of the S history are padded to an even pitch for the
    16-byte aligned bulk copies) and not multiples of the 64-day sweep."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = load_dataset(name)
    M = Model(ds, period=period, substeps=2)
    theta = np.array(load_golden(name)["theta"])
    logl, logp, status = M.calclogpost_batch(theta)
    ref = oracle.evaluate(ds["flavour"], ds, theta, period, 2, n_threads=8)
    assert np.array_equal(status, ref["status"])
    cond = oracle_conditioning(oracle, ds["flavour"], ds, theta, period, 2, ref["logl"])
    # the golden thetas include box-uniform draws; at 2 sub-steps/day one of the 49 is ill-conditioned
    # beyond what the jitter probe sees (same miss at the even period 300), so: at most one outlier,
    # and the well-conditioned family (init + near-mode rows) strictly.
    within = logl_within_bar(logl, ref["logl"], cond, TOL_LOGL)
    assert (~within).sum() <= 1, np.nonzero(~within)[0]
    assert rel_err(logl[:1], ref["logl"][:1]).max() <= TOL_LOGL
    # an odd period changes only the history pitch, not the arithmetic: bit-identical to the next even
    # period when the data window ends before either (age dataset; the simple datasets' window reaches the
    # period end for late offsets, where the period legitimately enters the result)
    M2 = Model(ds, period=period + 1, substeps=2)
    l2, p2, s2 = M2.calclogpost_batch(theta)
    okk = (status == 0) & (s2 == 0)
    if name == "A300":
        assert np.array_equal(logl[okk], l2[okk])
    M.close()


def test_async_submit_wait_matches_eval():
    """epi_eval_submit/epi_eval_wait (double-buffered host batches) return exactly what epi_eval returns,
    for both layouts, ragged sizes, interleaved slots and a workspace that grows between submissions."""
    from epi_mcmc_b200 import _capi
    M = _model("A300", 2)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(11)
    batches = [np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.05 * (mx - mn), mn, mx) for n in (257, 1000, 33, 4096, 1)]
    want = [M.calclogpost_batch(b) for b in batches]
    M2 = _model("A300", 2)
    got = [None] * len(batches)
    M2.calclogpost_submit(0, batches[0])
    for k in range(1, len(batches)):
        lay = "soa" if k % 2 else "aos"
        M2.calclogpost_submit(k % 2, batches[k].T.copy() if lay == "soa" else batches[k], layout=lay)
        got[k - 1] = M2.calclogpost_wait((k - 1) % 2)
    got[-1] = M2.calclogpost_wait((len(batches) - 1) % 2)
    for w, g in zip(want, got):
        for a, b in zip(w, g):
            assert np.array_equal(a, b, equal_nan=True)
    # call-sequence errors
    with pytest.raises(_capi.EpiError):
        M2.calclogpost_wait(0)
    M2.calclogpost_submit(1, batches[2])
    with pytest.raises(_capi.EpiError):
        M2.calclogpost_submit(1, batches[2])
    M2.calclogpost_wait(1)
    M.close()
    M2.close()


@pytest.mark.parametrize("name", ["A300", "I290", "S120", "NE120"])
def test_ext_trajectories_match_oracle(name, oracle):
    """epi_trajectories_ext: E, I, R and the Re/Rt output variables of derivs() at the integer output
    times (model-agen.cpp:159-163) — incl. invalid-offset draws (pass-1 matrix recycled), batches below
    and above the two-lanes-per-chain launch threshold, and a horizon other than FitTotalPeriod."""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset(name)
    M = _model(name, 4)
    g = load_golden(name)
    theta = np.array(g["theta"])[:20]
    model_th = M.transformParams(theta)
    nb = len(M.series_names)
    for P in (ds["period"], ds["period"] + 33):
        series, off, pad = M.calculateModel_batch(model_th, P, ext=True)
        basic, off_b, _ = M.calculateModel_batch(model_th, P)
        assert series.shape[1] == len(M.series_names_ext)
        assert np.array_equal(series[:, :nb], basic, equal_nan=True) and np.array_equal(off, off_b)
        for i in range(len(theta)):
            ref, _, roff, rpad = oracle.trajectory(ds["flavour"], ds, model_th[i], ds["period"], 4, calc_period=P, ext=True)
            assert off[i] == roff and pad[i] == rpad
            got, want = series[i, nb:], ref[nb:]
            # states: relative to the population; diagnostics (ratios of states): relative, except where the
            # denominators are still ~0 persons and the ratio amplifies the integrator's last bits
            n_state = 6 if M.is_age else 3
            bar_s = 1e-10 * np.maximum(np.abs(want[:n_state]), 1e-6 * ds["N"])
            assert (np.abs(got[:n_state] - want[:n_state]) <= bar_s).all()
            with np.errstate(invalid="ignore"):
                err = np.abs(got[n_state:] - want[n_state:]) / (1e-9 * np.abs(want[n_state:]) + 1e-12)
            err = np.where(np.isnan(got[n_state:]) & np.isnan(want[n_state:]), 0.0, err)
            assert not np.isnan(err).any() and err.max() <= 1.0, (i, err.max())
    st = M.calculateModel(model_th[0])
    assert "Re" in st and len(st["Re"]) == ds["period"]
    M.close()


def test_quantile_data_of_re(oracle):
    """quantileData(data_sample, function(state, params) state$Re, ...) (analyzeMCMC.R:265) on the device"""
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset("A300")
    M = _model("A300", 4)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(9)
    n, period, startoffset = 150, 100, 0
    sample = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.04 * (mx - mn), mn, mx)
    probs = [0.05, 0.5, 0.95]
    got = M.quantileData(sample, "Re", startoffset, period, probs)
    simperiod = 3 * period
    data = np.full((simperiod, n), np.nan)
    k = M.series_names_ext.index("Re")
    for i in range(n):
        ser, _, off, pad = oracle.trajectory("age2q", ds, sample[i], ds["period"], 4, calc_period=simperiod, ext=True)
        if off == 10000:
            continue
        vec = np.zeros(pad + simperiod)          # state$Re is pre-filled with 0 (model-age-groups2q.R:136)
        vec[pad:] = ser[k]
        take = vec[off - startoffset - 1:simperiod]
        data[:len(take), i] = take
    ref = np.nanquantile(data[:period], probs, axis=1).T
    assert np.abs(got - ref).max() <= 1e-9 * np.nanmax(np.abs(ref))
    M.close()


def test_two_devices_in_one_process():
    """Per-device launch state (opt-in shared-memory attribute, SM count): a second handle on another GPU of the
    same process must work, also for a period whose shared-memory footprint needs the opt-in."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from epi_mcmc_b200.model import Model, load_dataset
    ds = load_dataset("A365")
    th = np.array(load_golden("A365")["theta"])
    A = Model(ds, substeps=1, device=0)
    B = Model(ds, substeps=1, device=1)
    la, pa, sa = A.calclogpost_batch(th)
    lb, pb, sb = B.calclogpost_batch(th)
    assert np.array_equal(la, lb, equal_nan=True) and np.array_equal(pa, pb) and np.array_equal(sa, sb)
    A.close()
    B.close()


@pytest.mark.parametrize("name,m", [("A300", 4), ("A365", 1), ("I290", 2), ("S365", 2), ("NE120", 4), ("DE", 4)])
def test_window_limited_sweep_is_bit_identical(name, m, monkeypatch):
    """The scoring path integrates pass 2 only up to the last data-window day and sweeps only the window days
    (+ one predecessor).  Against sweeping the whole period (EPI_WINDOW_ONLY=0): the same offsets, statuses, NA
    pattern and log prior bit for bit, and the same per-day contributions — the two sweeps start on different days, so
    a lane of the warp sums a different subset of them and the totals may differ in the last bits (<= 1e-13
    relative) — incl. invalid offsets, clamped windows and NA draws of the golden theta sets."""
    from epi_mcmc_b200.model import Model, load_dataset
    theta = np.array(load_golden(name)["theta"])
    ds = load_dataset(name)
    monkeypatch.setenv("EPI_WINDOW_ONLY", "1")
    A = Model(ds, substeps=m)
    monkeypatch.setenv("EPI_WINDOW_ONLY", "0")
    B = Model(ds, substeps=m)
    la, pa, sa = A.calclogpost_batch(theta)
    lb, pb, sb = B.calclogpost_batch(theta)
    assert np.array_equal(np.isnan(la), np.isnan(lb)) and np.array_equal(pa, pb) and np.array_equal(sa, sb)
    assert np.allclose(la, lb, rtol=1e-13, atol=0, equal_nan=True)
    oa, _, ta = A.eval_detail(theta)
    ob, _, tb = B.eval_detail(theta)
    assert np.array_equal(oa, ob) and np.allclose(ta, tb, rtol=1e-13, atol=1e-13, equal_nan=True)
    A.close()
    B.close()


@pytest.mark.parametrize("name,m,period", [("S365", 4, None), ("S365", 1, 500), ("I290", 2, None), ("NE120", 4, 400)])
def test_two_round_pass1_is_bit_identical(name, m, period, monkeypatch, oracle):
    """A pass 1 that covers the whole period (simple / Ne / age-2i) is integrated for its first 128 days, and only
    chains whose threshold is not crossed there are redone over all days (retry round).  Against one round over
    all days (EPI_PASS1_CHUNK=0): identical bits for logl, offsets, statuses and trajectories — the theta sets
    include late crossings (day > 128) and chains that never cross."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = load_dataset(name)
    mn, mx, init = (np.array(v) for v in oracle.df_params(ds["flavour"], ds, ds["period"]))
    rng = np.random.default_rng(3)
    theta = np.vstack([np.array(load_golden(name)["theta"]), mn + rng.uniform(size=(1500, len(mn))) * (mx - mn)])
    P = period or ds["period"]
    monkeypatch.setenv("EPI_PASS1_CHUNK", "1")
    A = Model(ds, period=P, substeps=m)
    monkeypatch.setenv("EPI_PASS1_CHUNK", "0")
    B = Model(ds, period=P, substeps=m)
    la, pa, sa = A.calclogpost_batch(theta)
    lb, pb, sb = B.calclogpost_batch(theta)
    oa, qa, _ = A.eval_detail(theta)
    ob, qb, _ = B.eval_detail(theta)
    assert np.array_equal(oa, ob) and np.array_equal(qa, qb) and np.array_equal(sa, sb)
    assert np.array_equal(la, lb, equal_nan=True) and np.array_equal(pa, pb)
    first = oa + ds["lockdown_offset"] - qa - 1          # crossing day index within pass 1
    late = (oa != 10000) & (first >= 128)
    # the retry round really ran: chains that never cross and / or chains that cross behind the first chunk
    if name != "I290":  # (the age-2i box never delays the crossing beyond day ~70: its retry round stays empty)
        assert ((oa == 10000) | late).sum() > 0, first.max()
    if name == "S365":
        assert (oa == 10000).sum() > 0 and late.sum() > 0, first.max()
    ta, _, _ = A.calculateModel_batch(A.transformParams(theta[:64]), P + 50, ext=True)
    tb, _, _ = B.calculateModel_batch(B.transformParams(theta[:64]), P + 50, ext=True)
    assert np.array_equal(ta, tb, equal_nan=True)
    A.close()
    B.close()


@pytest.mark.parametrize("name,m", [("A300", 4), ("I290", 2), ("S365", 2), ("NE120", 4), ("DE", 1)])
def test_removed_compartment_fallback_is_bit_identical(name, m, monkeypatch):
    """The hot ODE kernels do not integrate the removed compartment R (it only enters the reference through the
    bounds guard, which cannot fire on R while S >= 1 at every stage input); chains where that is not certain
    are recomputed by the R-tracking instance.  Forcing EVERY chain through that instance
    (EPI_FORCE_R_FALLBACK=1) must not change a bit of logl, the terms, offsets, statuses or trajectories."""
    from epi_mcmc_b200.model import Model, load_dataset
    theta = np.array(load_golden(name)["theta"])
    ds = load_dataset(name)
    monkeypatch.setenv("EPI_FORCE_R_FALLBACK", "0")
    A = Model(ds, substeps=m)
    monkeypatch.setenv("EPI_FORCE_R_FALLBACK", "1")
    B = Model(ds, substeps=m)
    la, pa, sa = A.calclogpost_batch(theta)
    lb, pb, sb = B.calclogpost_batch(theta)
    assert np.array_equal(la, lb, equal_nan=True) and np.array_equal(pa, pb) and np.array_equal(sa, sb)
    oa, qa, ta = A.eval_detail(theta)
    ob, qb, tb = B.eval_detail(theta)
    assert np.array_equal(oa, ob) and np.array_equal(qa, qb) and np.array_equal(ta, tb, equal_nan=True)
    xa, _, _ = A.calculateModel_batch(A.transformParams(theta[:32]), ds["period"] + 25)
    xb, _, _ = B.calculateModel_batch(B.transformParams(theta[:32]), ds["period"] + 25)
    assert np.array_equal(xa, xb, equal_nan=True)
    assert B.launch_count() > 0
    A.close()
    B.close()


@pytest.mark.parametrize("name", ["A300", "I290"])
def test_both_ode_kernels_produce_the_same_bits(name):
    """Small batches of the age flavours use two lanes per chain (k_ode_age2), large ones a thread per chain
    (k_ode): the same theta must give the same bits through either."""
    M = _model(name, 4)
    theta = np.array(load_golden(name)["theta"])
    small = M.calclogpost_batch(theta)
    reps = 45000 // len(theta) + 1          # > 176 chains per SM: thread-per-chain launch
    big = M.calclogpost_batch(np.tile(theta, (reps, 1)))
    for a, b in zip(small, big):
        assert np.array_equal(a, b[:len(theta)], equal_nan=True)
        assert np.array_equal(a, b[-len(theta):], equal_nan=True)
    M.close()


@pytest.mark.parametrize("name,m", [("A300", 4), ("A300", 1), ("A300", 3), ("I290", 4), ("A365", 2)])
def test_ode_lane_mappings_produce_the_same_bits(name, m, monkeypatch):
    """EPI_ODE_LANES forces one, two or four lanes per chain for the ODE passes of the age flavours (k_ode,
    k_ode_age2, k_ode_age4).  Golden theta, box-uniform theta (guard slow path, invalid offsets, unsorted
    breakpoints) and a ragged batch size must give identical bits, offsets and statuses through all three."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = load_dataset(name)
    theta = np.array(load_golden(name)["theta"])
    rng = np.random.default_rng(11)
    outs = {}
    for lanes in (1, 2, 4):
        monkeypatch.setenv("EPI_ODE_LANES", str(lanes))
        M = Model(ds, substeps=m)
        mn, mx = M.df_params["min"], M.df_params["max"]
        if lanes == 1:
            box = mn + rng.uniform(size=(1003, M.d)) * (mx - mn)
            th = np.vstack([theta, box])
        logl, logp, status = M.calclogpost_batch(th)
        off, pad, terms = M.eval_detail(th)
        outs[lanes] = (logl, logp, status, off, pad, terms)
        M.close()
    for lanes in (2, 4):
        for a, b in zip(outs[1], outs[lanes]):
            assert np.array_equal(a, b, equal_nan=True), lanes


@pytest.mark.parametrize("ny,no,beta_scale", [(300.0, 120.0, 1.0), (4000.0, 1500.0, 1.0), (4000.0, 1500.0, 4.0), (8.55e6, 2.95e6, 4.0)])
def test_two_lane_kernel_under_stress(ny, no, beta_scale, monkeypatch):
    """The two-lane kernel only accumulates its maybe-out-of-bounds flags and votes once per day (the checked loop
    redoes a day some chain raised one in).  Small populations (S runs through [1, 4) and below 1: the R-tracking
    fallback), explosive epidemics (E, I beyond N/2) and the usual case must give the bits, offsets, statuses and
    fallback counts of the thread-per-chain kernel, which decides every step on its own."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = dict(load_dataset("A300"))
    ds["y_N"], ds["o_N"], ds["N"] = ny, no, ny + no
    outs = {}
    for lanes in (1, 2):
        monkeypatch.setenv("EPI_ODE_LANES", str(lanes))
        M = Model(ds, substeps=4)
        mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
        rng = np.random.default_rng(3)
        theta = np.clip(init + (rng.uniform(size=(777, 45)) - 0.5) * 0.3 * (mx - mn), mn, mx)
        names = list(M.fit_paramnames)
        for k, nm in enumerate(names):
            if nm.startswith("beta") or nm.startswith("y.beta") or nm.startswith("o.beta") or nm.startswith("yo.beta"):
                theta[:, k] = np.minimum(theta[:, k] * beta_scale, mx[k])
        if ny < 1e5:
            theta[:, 17] = rng.uniform(0.0, 0.5, size=len(theta))   # t0_morts: a threshold a small epidemic can cross
        series, off, pad = M.calculateModel_batch(theta[:256])
        outs[lanes] = M.calclogpost_batch(theta) + (M.fallback_count(), series, off, pad)
        M.close()
    assert np.isfinite(outs[1][0]).any()
    for a, b in zip(outs[1], outs[2]):
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("m", [4, 1, 3])
def test_two_lane_kernel_of_the_waning_immunity_flavour(m, monkeypatch):
    """age-2x on two lanes per chain (five states per lane: R is integrated and flows back into S, the younger group's
    effective-susceptible cap, the history rows hold the incidence) against thread per chain: golden theta,
    box-uniform theta and a ragged batch must give identical bits, offsets, statuses and trajectories."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = load_dataset("X300")
    theta = np.array(load_golden("X300")["theta"])
    rng = np.random.default_rng(29)
    outs = {}
    for lanes in (1, 2):
        monkeypatch.setenv("EPI_ODE_LANES", str(lanes))
        M = Model(ds, substeps=m)
        mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
        if lanes == 1:
            box = mn + rng.uniform(size=(1003, M.d)) * (mx - mn)
            near = np.clip(init + (rng.uniform(size=(1000, M.d)) - 0.5) * 0.3 * (mx - mn), mn, mx)
            th = np.vstack([theta, box, near])
        logl, logp, status = M.calclogpost_batch(th)
        off, pad, terms = M.eval_detail(th)
        outs[lanes] = (logl, logp, status, off, pad, terms)
        M.close()
    assert np.isfinite(outs[1][0]).mean() > 0.3
    for a, b in zip(outs[1], outs[2]):
        assert np.array_equal(a, b, equal_nan=True)


def test_four_lane_kernel_through_the_r_fallback(monkeypatch):
    """Tiny population: S falls below 1, the four-lane kernel flags the chains and the R-tracking thread-per-chain
    instance recomputes them — same bits as the thread-per-chain hot path followed by the same fallback."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = dict(load_dataset("A300"))
    ds["y_N"], ds["o_N"], ds["N"] = 300.0, 120.0, 420.0
    outs = {}
    for lanes in (1, 4):
        monkeypatch.setenv("EPI_ODE_LANES", str(lanes))
        M = Model(ds, substeps=4)
        mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
        rng = np.random.default_rng(2)
        theta = np.clip(init + (rng.uniform(size=(256, 45)) - 0.5) * 0.3 * (mx - mn), mn, mx)
        theta[:, 17] = rng.uniform(0.0, 0.5, size=256)
        outs[lanes] = M.calclogpost_batch(theta) + (M.fallback_count(),)
        M.close()
    assert outs[1][3] > 0 and outs[1][3] == outs[4][3]
    for a, b in zip(outs[1][:3], outs[4][:3]):
        assert np.array_equal(a, b, equal_nan=True)


def test_tiny_population_exercises_the_r_fallback(oracle, monkeypatch):
    """With a population of a few hundred the epidemic burns through everybody: S falls below 1, the hot ODE path
    can no longer rule out the R part of the bounds guard, and the chains are recomputed by the R-tracking instance.
    Results must still match the oracle (which integrates R and evaluates the full guard) and the forced-fallback run."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = dict(load_dataset("A300"))
    ds["y_N"], ds["o_N"], ds["N"] = 300.0, 120.0, 420.0
    mn, mx, init = (np.array(v) for v in oracle.df_params("age2q", ds, 300))
    rng = np.random.default_rng(2)
    theta = np.clip(init + (rng.uniform(size=(256, 45)) - 0.5) * 0.3 * (mx - mn), mn, mx)
    theta[:, 17] = rng.uniform(0.0, 0.5, size=256)   # t0_morts: a threshold such a small epidemic can cross
    M = Model(ds, substeps=4)
    logl, logp, status = M.calclogpost_batch(theta)
    n_fb = M.fallback_count()
    assert n_fb > 0, "the scenario no longer reaches S < 1: pick a smaller population"
    off, pad, _ = M.eval_detail(theta)
    ref = oracle.evaluate("age2q", ds, theta, 300, 4, n_threads=8)
    mism = (off != ref["offset"]) | (pad != ref["padding"]) | (status != ref["status"])
    assert mism.sum() <= 2, int(mism.sum())
    ok = ~mism
    cond = oracle_conditioning(oracle, "age2q", ds, theta, 300, 4, ref["logl"])
    within = logl_within_bar(logl[ok], ref["logl"][ok], cond[ok], TOL_LOGL)
    assert within.mean() >= 0.98, (int((~within).sum()), int(ok.sum()))
    monkeypatch.setenv("EPI_FORCE_R_FALLBACK", "1")
    B = Model(ds, substeps=4)
    lb, pb, sb = B.calclogpost_batch(theta)
    assert B.fallback_count() == len(theta)
    assert np.array_equal(logl, lb, equal_nan=True) and np.array_equal(status, sb)
    M.close()
    B.close()


@pytest.mark.parametrize("name,m", [("A300", 3), ("A300", 8), ("I290", 3), ("S120", 3), ("NE120", 7), ("S365", 8)])
def test_other_substep_counts(name, m, oracle):
    """sub-step counts whose step 1/m is not a power of two (m = 3, 7: the sub-step boundaries are rounded) and a
    finer grid (m = 8): same bar as the golden test — integer outputs exact, logl within the conditioning bar,
    the model file's init strictly."""
    from epi_mcmc_b200.model import Model, load_dataset
    ds = load_dataset(name)
    M = Model(ds, substeps=m)
    theta = np.array(load_golden(name)["theta"])
    logl, logp, status = M.calclogpost_batch(theta)
    off, pad, _ = M.eval_detail(theta)
    ref = oracle.evaluate(ds["flavour"], ds, theta, ds["period"], m, n_threads=8)
    mism = (off != ref["offset"]) | (pad != ref["padding"]) | (status != ref["status"])
    assert mism.sum() <= 1, int(mism.sum())
    ok = ~mism
    cond = oracle_conditioning(oracle, ds["flavour"], ds, theta, ds["period"], m, ref["logl"])
    within = logl_within_bar(logl[ok], ref["logl"][ok], cond[ok], TOL_LOGL)
    assert (~within).sum() <= 1, [(i, logl[ok][i], ref["logl"][ok][i], cond[ok][i]) for i in np.nonzero(~within)[0]]
    assert rel_err(logp, ref["logp"]).max() <= TOL_LOGP
    assert rel_err(logl[:1] + logp[:1], ref["logl"][:1] + ref["logp"][:1]).max() <= TOL_LOGL
    M.close()


def _box_jitter(M, n, seed, width):
    rng = np.random.default_rng(seed)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    return np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * width * (mx - mn), mn, mx)


@pytest.mark.parametrize("name,n,m", [("A300", 65536, 4), ("S365", 1 << 20, 4), ("NE120", 4096, 4)])
def test_full_size_properties(name, n, m, oracle):
    """BASELINE.json's full sizes (65,536 age chains; 2^20 simple proposals over 365 days; 4,096 Ne chains), checked
    through size-independent properties: (1) a proposal's result does not depend on where in the batch it sits or
    how big the batch is — the whole batch reversed, and slices of 4,099 (not a multiple of any tile), reproduce the
    same bits; (2) duplicated proposals give duplicated bits; (3) a seeded sample of 192 chains of the big batch
    meets the oracle bar."""
    M = _model(name, m)
    theta = _box_jitter(M, n, 11, 0.2)
    theta[n // 2:n // 2 + 64] = theta[:64]                         # (2)
    l, p, s = M.calclogpost_batch(theta)
    assert np.array_equal(l[n // 2:n // 2 + 64], l[:64], equal_nan=True) and np.array_equal(p[n // 2:n // 2 + 64], p[:64])
    lr, pr, sr = M.calclogpost_batch(theta[::-1].copy())           # (1) position
    assert np.array_equal(lr[::-1], l, equal_nan=True) and np.array_equal(pr[::-1], p) and np.array_equal(sr[::-1], s)
    step = 4099
    for a in range(0, min(n, 8 * step), step):                     # (1) batch size: the pair kernel / thread kernel switch
        ls, ps, ss = M.calclogpost_batch(theta[a:a + step])
        assert np.array_equal(ls, l[a:a + step], equal_nan=True) and np.array_equal(ps, p[a:a + step])
        assert np.array_equal(ss, s[a:a + step])
    pick = np.random.default_rng(5).choice(n, 192, replace=False)  # (3)
    ds, P = M.data, M.data["period"]
    o = oracle.evaluate(M.flavour, ds, theta[pick], P, m, n_threads=8)
    same = s[pick] == o["status"]                                  # a threshold touched within rounding: <= 1 (see above)
    assert (~same).sum() <= 1
    assert np.all(rel_err(p[pick], o["logp"]) <= TOL_LOGP)
    cond = oracle_conditioning(oracle, M.flavour, ds, theta[pick], P, m, o["logl"])
    ok = logl_within_bar(l[pick], o["logl"], cond, TOL_LOGL) | ~same
    assert ok.all(), (np.flatnonzero(~ok), l[pick][~ok], o["logl"][~ok])
    assert (s[pick] == 0).sum() > 96
    M.close()
