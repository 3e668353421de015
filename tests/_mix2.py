"""one-off diagnostic: burn-in strategies for the DE-MC population (not a test)"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler
M = Model(load_dataset("A300"), substeps=4)
n = 16384
mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
keys = [M.fit_paramnames.index(k) for k in ("betay6", "betao7", "ifrred", "y.DL", "t0_morts", "y.CR")]
def report(tag, S, it):
    th, ll, lp = S.get()
    sd = th.std(axis=0) / (mx - mn)
    lpost = ll + lp
    print(tag, it, "acc %.3f" % S.stats()["accept_rate"], "mean %.2f med %.2f p05 %.2f max %.2f" % (np.mean(lpost), np.median(lpost), np.percentile(lpost, 5), np.max(lpost)),
          "sd/box", np.round(sd[keys], 4), flush=True)
for jit, tune in ((0.02, False), (0.02, True), (0.10, True)):
    rng = np.random.default_rng(3)
    th0 = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * jit * (mx - mn), mn, mx)
    S = Sampler(M, n, kind="demc", seed=1, theta0=th0)
    S.adapt(True)
    for blk in range(1, 41):
        for _ in range(10):
            S.run(10); S.refresh_population()
            if tune and blk <= 20: S.tune_scale(target=0.2)
        if blk == 20: S.adapt(False)
        if blk in (1, 2, 5, 10, 15, 20, 30, 40): report("jitter %.2f tune %d scale %.2f" % (jit, tune, S.scale), S, blk * 100)
