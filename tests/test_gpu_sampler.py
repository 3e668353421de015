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


def test_rwm_and_demc_agree_on_the_posterior():
    """two different samplers, one target (simple model, synthetic S120): pooled posterior means
    agree within Monte-Carlo error, and the truth used to generate the data is covered."""
    from epi_mcmc_b200.sampler import run_mcmc
    M = _model("S120", 2)
    r1 = run_mcmc(M, burnin=600, samples=300, chains=1024, kind="demc", seed=1, thin=3)
    r2 = run_mcmc(M, burnin=1500, samples=600, chains=1024, kind="rwm", seed=2, thin=6)
    d1 = r1["draws"].reshape(-1, M.d)
    d2 = r2["draws"].reshape(-1, M.d)
    sd = d1.std(axis=0)
    z = np.abs(d1.mean(axis=0) - d2.mean(axis=0)) / sd
    assert (z < 0.35).all(), z
    q1, q2 = np.quantile(d1, [0.1, 0.9], axis=0), np.quantile(d2, [0.1, 0.9], axis=0)
    assert (np.abs(q1 - q2) / sd < 0.5).all()
    truth = M.df_params["init"]
    lo, hi = np.quantile(d1, [0.001, 0.999], axis=0)
    assert ((truth >= lo) & (truth <= hi))[:7].all()
    assert 0.05 < r1["diagnostics"]["accept_rate"] < 0.8
    assert min(r1["diagnostics"]["ess"].values()) > 100
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
