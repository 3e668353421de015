"""one-off diagnostic (not a test): is the DE-MC population stationary after burn-in?"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler
M = Model(load_dataset("A300"), substeps=4)
n = 16384
S = Sampler(M, n, kind="demc", seed=1)
mn, mx = M.df_params["min"], M.df_params["max"]
keys = [M.fit_paramnames.index(k) for k in ("betay6", "betao7", "ifrred", "y.DL", "t0_morts", "y.CR")]
S.adapt(True)
it = 0
for blk in range(60):
    for _ in range(10):
        S.run(10); S.refresh_population()
    it += 100
    if it == 1000: S.adapt(False)
    th, ll, lp = S.get()
    sd = th.std(axis=0) / (mx - mn)
    print(it, "acc %.3f" % S.stats()["accept_rate"], "mean lpost %.2f" % np.mean(ll + lp), "max %.2f" % np.max(ll + lp),
          "sd/box", np.round(sd[keys], 4), flush=True)
