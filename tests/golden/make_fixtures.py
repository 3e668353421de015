#!/usr/bin/env python
"""Generates (a) the named datasets under epi_mcmc_b200/datasets/ and (b) the
golden known-answer vectors under tests/golden/.

Run in the build container only (it reads /root/reference/R/data/ecdc/ecdc.csv
for the real-data sets and uses the CPU oracle to draw the synthetic observed
series).  The outputs are committed; nothing at test/bench time reads
/root/reference.

Workloads:
  S120, S365   synthetic simple SEIR, P = 120 / 365
  NE120        synthetic Ne flavour, P = 120
  A300, A365   synthetic two-age-group (model-age-groups2q), P = 300 / 365
  I290         synthetic two-age-group, earlier generation (model-age-groups2i,
               analyses/be/age/settings.R: P = 290, sizes 0.1 / 2 / 20 / 120)
  DE, SE       ECDC daily cases/deaths (R/data/ecdc/ecdc.csv, geoId DE / SE),
               loader logic of R/data/ecdc/data.R:75-105 with the lockdown
               dates fixed by hand (the Google mobility file is not in the tree)
"""
import csv
import datetime as dt
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

DS_DIR = os.path.join(ROOT, "epi_mcmc_b200", "datasets")
GOLD_DIR = os.path.join(ROOT, "tests", "golden")
ECDC = "/root/reference/R/data/ecdc/ecdc.csv"
M_DEFAULT = 4


def glogis(t, A, K, C, Q, B, M, v):
    """R/data/be/data-age-all.R:5-7"""
    return A + (K - A) / ((C + Q * np.exp(-B * (t - M))) ** (1 / v))


def nb_draw(rng, mu, size):
    mu = np.maximum(mu, 1e-9)
    p = size / (size + mu)
    return rng.negative_binomial(size, p).astype(float)


def simple_dataset(P, flavour="simple"):
    n = P - 40
    d = dict(name=("S" if flavour == "simple" else "NE") + str(P), flavour=flavour, period=P,
             N=8.3e7, lockdown_offset=12, lockdown_transition_period=10,
             total_deaths_at_lockdown=5.0, hosp_nbinom_size=15.0, mort_nbinom_size=30.0,
             dhospi=[0.0] * n, dmorti=[0.0] * n, n_dhospi=n, n_dmorti=n, n_dhosp=n, dmort_last=500.0)
    # init: R/models/model-simple-ode-drj-cmp.R:263 (+ Nef = 0.7 for the Ne flavour)
    init = [2.9, 0.9, 3, 3, np.log(0.02), 10, 20, 5.0] + ([0.7] if flavour == "ne" else [])
    model = list(init)
    model[4] = float(np.exp(init[4]))
    series, _, off, pad = oracle.trajectory(flavour, d, model, P, M_DEFAULT)
    assert off != 10000, "init does not cross the lockdown threshold"
    rng = np.random.default_rng(20200310)
    # data day i <-> model index data_offset + i - 1 (calclogl, :206-243); series[] starts at pad+1
    idx = off - 1 - pad + np.arange(n)
    assert idx.min() >= 0 and idx.max() < P, (off, pad, idx.min(), idx.max())
    d["dhospi"] = nb_draw(rng, np.maximum(0.1, series[1][idx]), 15.0).tolist()
    d["dmorti"] = nb_draw(rng, np.maximum(0.1, series[2][idx]), 30.0).tolist()
    d["dmort_last"] = float(np.sum(d["dmorti"]))
    d["provenance"] = ("synthetic: oracle (RK4 m=4) at the model file's init, NB noise sizes 15/30, "
                       "numpy default_rng(20200310); SURVEY.md §8d")
    return d


def age_dataset(P):
    n = P - 60
    nr = n + 200
    vs = np.arange(1, nr + 1, dtype=float)
    g = glogis(vs, 0, 1, 1, 0.5, 0.1, 200, 0.5)
    gtrimp = glogis(vs, 0, 1, 1, 0.5, 0.1, 113, 0.5)
    d = dict(name="A" + str(P), flavour="age2q", period=P, N=11.5e6, y_N=8.55e6, o_N=2.95e6,
             lockdown_offset=2, lockdown_transition_period=10, total_deaths_at_lockdown=5.0,
             d3=88, d4=118, d5=139, d6=160, d7=195, d8=220,
             d_reliable_cases=205, d_reliable_hosp=83, d_hosp_o1=104, d_hosp_o2=188,
             n_dhospi=n, n_dhosp=n, n_dmorti=n, n_dcasei=n, n_rate=nr,
             # y.ifr deviates from SURVEY §8d's 6e-4 (which pins pmin(1, y.CR*y.ifr) at its
             # cap for the whole box around init = 2000): 3e-4 keeps the case rate identifiable
             y_ifr=(3e-4 * (1 - 0.5 * g)).tolist(), o_ifr=[0.04] * nr, y_hr=[0.01] * nr, o_hr=[0.06] * nr,
             g=g.tolist(), gtrimp=gtrimp.tolist(), o_wdmorti=[1.0] * n,
             case_nbinom_size1=0.01, case_nbinom_sizey2=1.0, case_nbinom_sizeo2=1.0,
             hosp_nbinom_size1=20.0, hosp_nbinom_size2=40.0, mort_nbinom_size=120.0, hosp_last7=3.0,
             dhospi=[0.0] * n, y_dcasei=[0.0] * n, o_dcasei=[0.0] * n, y_dmorti=[0.0] * n,
             o_dmorti=[0.0] * n, dmort_last=20000.0)
    _, _, init = oracle.df_params("age2q", d, P)
    series, _, off, pad = oracle.trajectory("age2q", d, init, P, M_DEFAULT)
    assert off != 10000, "init does not cross t0_morts within 60 days"
    idx = off - 1 - pad + np.arange(n)
    assert idx.min() >= 0 and idx.max() < P, (off, pad, idx.min(), idx.max())
    rng = np.random.default_rng(20200310)
    ycasei, ocasei, yhospi, ohospi, ydeadi, odeadi = (series[q][idx] for q in range(2, 8))
    r = d["d_reliable_cases"]
    # observed cases: regime 1 drawn with size 1 (the fit's 0.01 only says "unreliable"), regime 2 per settings
    sz_case = np.where(np.arange(1, n + 1) < r, 1.0, 1.0)
    d["y_dcasei"] = nb_draw(rng, np.maximum(0.1, ycasei), sz_case).tolist()
    d["o_dcasei"] = nb_draw(rng, np.maximum(0.1, ocasei), sz_case).tolist()
    rh = d["d_reliable_hosp"]
    sz_h = np.where(np.arange(1, n + 1) < rh, 20.0, 40.0)
    d["dhospi"] = nb_draw(rng, np.maximum(0.1, yhospi + ohospi), sz_h).tolist()
    d["y_dmorti"] = nb_draw(rng, np.maximum(0.001, ydeadi), 240.0).tolist()
    d["o_dmorti"] = nb_draw(rng, np.maximum(0.001, odeadi), 120.0).tolist()
    d["dmort_last"] = float(np.sum(d["y_dmorti"]) + np.sum(d["o_dmorti"]))
    d["provenance"] = ("synthetic: oracle (RK4 m=4) at model-age-groups2q.R init, NB noise with the "
                       "analyses/be/age2q/settings.R sizes, numpy default_rng(20200310); constants per SURVEY.md §8d")
    return d


def age2i_dataset(P):
    n = P - 60
    nr = n + 200
    vs = np.arange(1, nr + 1, dtype=float)
    g = glogis(vs, 0, 1, 1, 0.5, 0.1, 200, 0.5)
    gtrimp = glogis(vs, 0, 1, 1, 0.5, 0.1, 113, 0.5)
    d = dict(name="I" + str(P), flavour="age2i", period=P, N=11.5e6, y_N=8.55e6, o_N=2.95e6,
             lockdown_offset=2, lockdown_transition_period=10, total_deaths_at_lockdown=5.0,
             d3=88, d4=118, d5=139, d6=160, d7=195, d8=220, d_reliable_cases=205,
             n_dhospi=n, n_dhosp=n, n_dmorti=n, n_dcasei=n, n_rate=nr,
             # this generation has no fyifr / ifrred parameters: the time dependence lives in the data vectors
             y_ifr=(3e-4 * (1 - 0.5 * g)).tolist(), o_ifr=(0.04 * (1 - 0.3 * gtrimp)).tolist(),
             y_hr=[0.01] * nr, o_hr=(0.06 * (1 - 0.2 * g)).tolist(),
             case_nbinom_size1=0.1, case_nbinom_sizey2=2.0, case_nbinom_sizeo2=2.0,
             hosp_nbinom_size=20.0, mort_nbinom_size=120.0,
             dhospi=[0.0] * n, y_dcasei=[0.0] * n, o_dcasei=[0.0] * n, y_dmorti=[0.0] * n,
             o_dmorti=[0.0] * n, dmort_last=20000.0)
    _, _, init = oracle.df_params("age2i", d, P)
    series, _, off, pad = oracle.trajectory("age2i", d, init, P, M_DEFAULT)
    assert off != 10000, "init does not cross t0_morts"
    idx = off - 1 - pad + np.arange(n)
    assert idx.min() >= 0 and idx.max() < P, (off, pad, idx.min(), idx.max())
    rng = np.random.default_rng(20200310)
    ycasei, ocasei, yhospi, ohospi, ydeadi, odeadi = (series[q][idx] for q in range(2, 8))
    d["y_dcasei"] = nb_draw(rng, np.maximum(0.1, ycasei), 2.0).tolist()
    d["o_dcasei"] = nb_draw(rng, np.maximum(0.1, ocasei), 2.0).tolist()
    d["dhospi"] = nb_draw(rng, np.maximum(0.001, yhospi + ohospi), 20.0).tolist()
    d["y_dmorti"] = nb_draw(rng, np.maximum(0.001, ydeadi), 120.0).tolist()
    d["o_dmorti"] = nb_draw(rng, np.maximum(0.001, odeadi), 120.0).tolist()
    d["dmort_last"] = float(np.sum(d["y_dmorti"]) + np.sum(d["o_dmorti"]))
    d["provenance"] = ("synthetic: oracle (RK4 m=4) at model-age-groups2i.R init, NB noise with the "
                       "analyses/be/age/settings.R sizes, numpy default_rng(20200310)")
    return d


def age2x_dataset(P):
    """the later generation (model-age-groups2x.R + model-ager.cpp).  No settings.R of the snapshot points at it, so
    its extra globals (d9, d10, d12, duncertain, d.symp.cases, d.all.cases, symp.cases.factor) are chosen here to be
    consistent with the file's own init vector (t3o = 92, t4o = 29, t6o = 28, t7o = 32, t10o = 245)."""
    n = P - 60
    nr = n + 200
    vs = np.arange(1, nr + 1, dtype=float)
    g = glogis(vs, 0, 1, 1, 0.5, 0.1, 200, 0.5)
    gtrimp = glogis(vs, 0, 1, 1, 0.5, 0.1, 113, 0.5)
    d = dict(name="X" + str(P), flavour="age2x", period=P, N=11.5e6, y_N=8.55e6, o_N=2.95e6,
             lockdown_offset=2, lockdown_transition_period=10, total_deaths_at_lockdown=5.0,
             d3=104, d4=133, d5=154, d6=182, d7=216, d8=225, d9=232, d10=245, d12=262, duncertain=280,
             d_symp_cases=228, d_all_cases=250, symp_cases_factor=0.8,
             d_reliable_cases=205, d_reliable_hosp=83, d_hosp_o1=104, d_hosp_o2=188,
             n_dhospi=n, n_dhosp=n, n_dmorti=n, n_dcasei=n, n_rate=nr,
             y_ifr=(3e-4 * (1 - 0.5 * g)).tolist(), o_ifr=[0.04] * nr, y_hr=[0.01] * nr, o_hr=[0.06] * nr,
             g=g.tolist(), gtrimp=gtrimp.tolist(), o_wdmorti=[1.0] * n,
             case_nbinom_size1=0.01, case_nbinom_sizey2=1.0, case_nbinom_sizeo2=1.0,
             hosp_nbinom_size1=20.0, hosp_nbinom_size2=40.0, mort_nbinom_size=120.0, hosp_last7=3.0,
             dhospi=[0.0] * n, y_dcasei=[0.0] * n, o_dcasei=[0.0] * n, y_dmorti=[0.0] * n,
             o_dmorti=[0.0] * n, dmort_last=20000.0)
    _, _, init = oracle.df_params("age2x", d, P)
    series, _, off, pad = oracle.trajectory("age2x", d, init, P, M_DEFAULT)
    assert off != 10000, "init does not cross t0_morts within 60 days"
    idx = off - 1 - pad + np.arange(n)
    assert idx.min() >= 0 and idx.max() < P, (off, pad, idx.min(), idx.max())
    rng = np.random.default_rng(20200310)
    ycasei, ocasei, yhospi, ohospi, ydeadi, odeadi = (series[q][idx] for q in range(2, 8))
    d["y_dcasei"] = nb_draw(rng, np.maximum(0.1, ycasei), 1.0).tolist()
    d["o_dcasei"] = nb_draw(rng, np.maximum(0.1, ocasei), 1.0).tolist()
    rh = d["d_reliable_hosp"]
    sz_h = np.where(np.arange(1, n + 1) < rh, 20.0, 40.0)
    d["dhospi"] = nb_draw(rng, np.maximum(0.1, yhospi + ohospi), sz_h).tolist()
    d["y_dmorti"] = nb_draw(rng, np.maximum(0.001, ydeadi), 600.0).tolist()
    d["o_dmorti"] = nb_draw(rng, np.maximum(0.001, odeadi), 120.0).tolist()
    d["dmort_last"] = float(np.sum(d["y_dmorti"]) + np.sum(d["o_dmorti"]))
    d["provenance"] = ("synthetic: oracle (RK4 m=4) at model-age-groups2x.R init, NB noise with the age2q settings' sizes, "
                       "numpy default_rng(20200310); the generation's extra globals are chosen to match its init vector")
    return d


def ecdc_dataset(geo, d1, transition, data_days, flavour):
    rows = [r for r in csv.DictReader(open(ECDC)) if r["geoId"] == geo]
    for r in rows:
        r["date"] = dt.date(int(r["year"]), int(r["month"]), int(r["day"]))
    d2 = d1 + dt.timedelta(days=transition)
    rows = [r for r in rows if r["date"] <= d2 + dt.timedelta(days=data_days)]  # data.R:79
    cases = np.array([float(r["cases"]) for r in rows])
    deaths = np.array([float(r["deaths"]) for r in rows])
    nzc = np.nonzero(cases > 0)[0] + 1
    nzm = np.nonzero(deaths > 0)[0] + 1
    starti = min(nzm[-1] + 30, nzc[-1])  # data.R:81-87 (rows are newest first)
    starti = min(starti, len(rows))
    dstart = rows[starti - 1]["date"]
    dhospi = np.maximum(0, cases[:starti][::-1])
    dmorti = np.maximum(0, deaths[:starti][::-1])
    lo = (d1 - dstart).days
    dmort = np.cumsum(dmorti)
    n = len(dhospi)
    return dict(name=geo + ("_NE" if flavour == "ne" else ""), flavour=flavour, period=130,
                N=float(rows[0]["popData2018"]), lockdown_offset=lo, lockdown_transition_period=transition,
                total_deaths_at_lockdown=float(max(1, dmort[lo - 1])), hosp_nbinom_size=15.0,
                mort_nbinom_size=30.0, dhospi=dhospi.tolist(), dmorti=dmorti.tolist(), n_dhospi=n,
                n_dmorti=n, n_dhosp=n, dmort_last=float(dmort[-1]), dstartdate=str(dstart),
                provenance=f"R/data/ecdc/ecdc.csv geoId {geo} through the logic of R/data/ecdc/data.R:75-105; "
                           f"d1={d1} and transition={transition} fixed by hand (mobility file absent); "
                           "sizes/period from analyses/de/simple/settings.R:29-32")


def conditioning(fl, ds, theta, P, m, logl):
    """see oracle.conditioning: the reference algorithm's own rounding sensitivity at theta"""
    return oracle.conditioning(fl, ds, theta, P, m, logl)


def golden(ds, n_random, seed):
    """theta set = init + box-interior draws; values from the oracle at m = 1 and m = 4."""
    fl, P = ds["flavour"], ds["period"]
    mn, mx, init = oracle.df_params(fl, ds, P)
    rng = np.random.default_rng(seed)
    u = rng.uniform(size=(n_random, len(mn)))
    theta = np.vstack([init, mn + u * (mx - mn)])
    # a second family close to init (valid data_offset much more often than box-uniform draws)
    near = init + (rng.uniform(size=(n_random, len(mn))) - 0.5) * 0.2 * (mx - mn)
    near = np.clip(near, mn, mx)
    theta = np.vstack([theta, near])
    out = dict(dataset=ds["name"], flavour=fl, period=P, theta=theta.tolist(), results={})
    for m in (1, 4):
        r = oracle.evaluate(fl, ds, theta, P, m)
        cond = conditioning(fl, ds, theta, P, m, r["logl"])
        out["results"][str(m)] = dict(
            cond=cond.tolist(),
            logl=[None if np.isnan(v) else v for v in r["logl"].tolist()],
            logp=r["logp"].tolist(), status=r["status"].tolist(), offset=r["offset"].tolist(),
            padding=r["padding"].tolist(),
            terms=[[None if np.isnan(v) else v for v in row] for row in r["terms"].tolist()])
    # trajectories at init (model space) for the calculateModel parity test
    model = init.copy()
    if fl not in ("age2q", "age2i", "age2x"):
        model[4] = np.exp(init[4])
    series, states, off, pad = oracle.trajectory(fl, ds, model, P, 4)
    out["init_trajectory"] = dict(substeps=4, offset=off, padding=pad,
                                  head=series[:, :5].tolist(), tail=series[:, -5:].tolist(),
                                  checksum=[float(np.sum(s)) for s in series])
    return out


def main():
    os.makedirs(DS_DIR, exist_ok=True)
    makers = {"S120": lambda: simple_dataset(120), "S365": lambda: simple_dataset(365),
              "NE120": lambda: simple_dataset(120, "ne"), "A300": lambda: age_dataset(300),
              "A365": lambda: age_dataset(365), "I290": lambda: age2i_dataset(290), "X300": lambda: age2x_dataset(300),
              "DE": lambda: ecdc_dataset("DE", dt.date(2020, 3, 13), 10, 60, "simple"),
              "SE_NE": lambda: ecdc_dataset("SE", dt.date(2020, 3, 12), 14, 60, "ne")}
    only = sys.argv[1:]  # optional: regenerate only the named sets
    sets = [mk() for name, mk in makers.items() if not only or name in only]
    for ds in sets:
        with open(os.path.join(DS_DIR, ds["name"] + ".json"), "w") as f:
            json.dump(ds, f)
        nrand = 24 if ds["flavour"] in ("age2q", "age2i", "age2x") else 32
        g = golden(ds, nrand, 1234)
        with open(os.path.join(GOLD_DIR, "golden_" + ds["name"] + ".json"), "w") as f:
            json.dump(g, f)
        r4 = g["results"]["4"]
        ok = sum(1 for s in r4["status"] if s == 0)
        print(ds["name"], "n_obs", ds["n_dhospi"], "golden thetas", len(g["theta"]), "status==0:", ok,
              "init logl", r4["logl"][0], "logp", r4["logp"][0], "offset", r4["offset"][0], "pad", r4["padding"][0])


if __name__ == "__main__":
    main()
