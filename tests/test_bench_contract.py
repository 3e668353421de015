"""The bench line contract, checked on committed lines (profiles/r2_bench_*.json were written by `python bench.py` /
`bench.py --impl reference` / torchrun on the B200 box) and on bench.py's own arithmetic helpers.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

import bench


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e")


OWN_LINES = [("r2_bench_n1_e.json", 1), ("r2_bench_n8_f.json", 8)]


@pytest.mark.parametrize("name,n_gpus", OWN_LINES)
def test_own_arm_line(name, n_gpus):
    j = _line(name)
    for k in BASE_KEYS + ("clocks", "gpu_launches", "roofline", "sustained", "secondary"):
        assert k in j, k
    assert j["metric"] == "log_posterior_evals_per_s" and j["unit"] == "evals/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == n_gpus and j["scaling"] == "strong" and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric
    c = j["config"]
    assert "workload" in c and "model" not in c
    # strong scaling: the configuration's 65,536 chains are sharded, value counts every one of them once per step
    assert c["chains_total"] == 65536 and c["chains_per_gpu"] * n_gpus == c["chains_total"]
    assert j["warmup"] >= 3 and j["gpu_launches"] >= 4 * j["steps"]
    assert j["value"] == pytest.approx(c["chains_total"] / (j["ms_per_step"] * 1e-3), rel=1e-6)
    e = j["e2e"]
    assert e["unit"] == j["unit"] and e["h2d_bytes_per_step"] == c["chains_per_gpu"] * c["params"] * 8
    assert e["d2h_bytes_per_step"] == c["chains_per_gpu"] * 20
    assert 0 < e["value"] < j["value"]                    # copies inside the timed region cost something
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0 < r["frac"] < 1
    # no launch is ever given a roofline fraction above 1 (the survey-formula figures are labelled as such)
    for k, v in j["roofline_all"].items():
        assert v["frac"] is None or 0 < v["frac"] < 1, (k, v["frac"])
    s = j["sustained"]
    assert s["steps"] * s["ms_per_step"] >= 990.0 and s["value"] == pytest.approx(j["value"], rel=0.05)
    cl = j["clocks"]
    assert cl["sm_mhz"] > 0.9 * cl["sm_max_mhz"]
    assert not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n_gpus == 1:
        b = j["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample", "r_probe", "ess_per_s_median"):
            assert k in b, k
        assert b["kind"] in ("port", "reference", "r") and b["cores"] >= 1
    else:
        assert j["weak"]["chains_per_gpu"] == 65536 and j["weak"]["value"] > j["value"]
    # the per-kernel event times add up to the step (launch gaps and the two idle fallback launches are the rest)
    assert sum(j["kernels"].values()) == pytest.approx(j["ms_per_step"], rel=0.03)
    # every other BASELINE config is observed by the same run
    got = {(x["workload"], x["substeps"]) for x in j["secondary"]}
    assert {("S120", 4), ("S365", 4), ("NE120", 4), ("A365", 4), ("A300", 1), ("I290", 4)} <= got
    for x in j["secondary"]:
        assert x["value"] == pytest.approx(x["chains_total"] / (x["ms_per_step"] * 1e-3), rel=1e-6)
    if "sampler" in j:
        sm = j["sampler"]
        assert 0 < sm["evals_per_s"] <= j["value"] * 1.05 and sm["replicate_populations"] >= 8
        k2 = {x["chains"]: x for x in sm["k2"]}
        assert set(k2) == {65536, 1 << 22}
        for x in k2.values():
            assert x["achieved_GBs"] == pytest.approx(x["algorithmic_bytes"] / (x["us_per_iter"] * 1e-6) / 1e9, rel=1e-9)
            assert 0 < x["frac"] < 1


def test_reference_arm_line():
    j = _line("r2_bench_ref_e.json")
    for k in BASE_KEYS + ("impl", "cpu_baseline"):
        assert k in j, k
    assert j["impl"] == "reference" and j["metric"] == "log_posterior_evals_per_s"
    own = _line("r2_bench_n1_e.json")
    assert j["config"] == own["config"] and j["unit"] == own["unit"] and j["scaling"] == own["scaling"]
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["cores"] >= 1
    assert "r_probe" in j["cpu_baseline"] and j["cpu_baseline"]["kind"] == ("r" if j["cpu_baseline"]["r_probe"]["usable"] else "port")


def test_f_alg_matches_the_survey_formula():
    """SURVEY.md 8d, age model at P = 300, m = 4: steps = m (P1 + P), ODE = steps (4 F_rhs + 14 S)"""
    fa = bench.f_alg("age2q", 300, 4, 31.4, 1215, 38)
    assert fa["ode"] == 4 * (60 + 300) * (4 * 124 + 14 * 10)
    assert fa["per_kernel"]["ode_pass2"] == 4 * 300 * (4 * 124 + 14 * 10) == 763200
    assert sum(fa["per_kernel"].values()) == pytest.approx(fa["total"])
    assert 1.1e6 < fa["total"] < 1.2e6
    fs = bench.f_alg("simple", 365, 4, 16.0, 600, 3)
    assert fs["ode"] == 4 * (365 + 365) * (4 * 72 + 14 * 4)


def test_bench_theta_clouds_stay_inside_the_box():
    from oracle import oracle as O
    from epi_mcmc_b200.model import load_dataset
    ds = load_dataset("A300")
    mn, mx, init = O.df_params("age2q", ds, 300, 4)
    th = bench.make_theta("A300", 2048, 7, mn, mx, init)
    assert th.shape == (2048, 45) and (th >= mn).all() and (th <= mx).all()
    assert np.abs(th - init).max() <= 0.005 * (mx - mn).max() + 1e-12
    with open(os.path.join(ROOT, "epi_mcmc_b200", "datasets", "A300_posterior_gauss.json")) as f:
        g = json.load(f)
    mu, cov = np.array(g["mean"]), np.array(g["cov"])
    assert mu.shape == (45,) and cov.shape == (45, 45) and np.allclose(cov, cov.T)
    assert (mu > mn).all() and (mu < mx).all() and np.linalg.eigvalsh(cov).min() > -1e-9
