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
for kind in ("rwm",):
    S = Sampler(M, n, kind=kind, seed=1)
    S.adapt(True)
    for blk in range(1, 41):
        for _ in range(10):
            S.run(10)
        if blk == 20: S.adapt(False)
        if blk in (1, 2, 5, 10, 15, 20, 30, 40): report(kind, S, blk * 100)
