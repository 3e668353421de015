"""diagnostic (not collected): SURVEY 8(f)-1 — quantileData over a posterior sample (libMCMC.R:151-172: the reference
calls calculateModel(params, 3 * period) per draw) on the GPU against the CPU oracle's calculateModel.
run: python tests/_traj_bench.py"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epi_mcmc_b200.model import Model, load_dataset
from oracle import oracle

oracle.build()
ds = load_dataset("A300")
M = Model(ds, substeps=4)
g = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "epi_mcmc_b200", "datasets",
                                "A300_posterior_gauss.json")))
rng = np.random.default_rng(1)
mn, mx = M.df_params["min"], M.df_params["max"]
for n in (1000, 8192):
    sample = np.clip(rng.multivariate_normal(np.array(g["mean"]), np.array(g["cov"]), size=n), mn, mx)
    for fun in (("y.hospi", "o.hospi"), "Re"):
        M.quantileData(sample[:64], fun, 0, 200, [0.05, 0.5, 0.95])          # warm-up (workspace allocation)
        t0 = time.perf_counter()
        q = M.quantileData(sample, fun, 0, 200, [0.05, 0.5, 0.95])
        dt = time.perf_counter() - t0
        print(f"quantileData {fun}: {n} draws x calculateModel(600 days) + 200 days x 3 quantiles: {dt * 1e3:.1f} ms "
              f"({n / dt:.0f} draws/s)", flush=True)
    t0 = time.perf_counter()
    series, off, pad = M.calculateModel_batch(M.transformParams(sample), period=900, ext=True)
    dt = time.perf_counter() - t0
    print(f"calculateModel_batch ext, {n} draws x 900 days x 18 series to the host: {dt * 1e3:.1f} ms ({n / dt:.0f} draws/s, "
          f"{series.nbytes / 1e6:.0f} MB)", flush=True)
k = 48
t0 = time.perf_counter()
for i in range(k):
    oracle.trajectory("age2q", ds, sample[i], 300, 4, calc_period=600, ext=True)
dt = time.perf_counter() - t0
print(f"CPU oracle calculateModel(600 days), one thread: {dt / k * 1e3:.2f} ms per draw ({k / dt:.0f} draws/s)")
