"""diagnostic (not collected): Monte-Carlo efficiency of the two on-device samplers on the age posterior, measured the
same way (independent replicate populations, ESS of the posterior-mean estimator from the spread of the replicate
means; growth of the ESS between the first half and the whole recorded window per evaluation).
run: python tests/_ess_compare.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler

M = Model(load_dataset("A300"), substeps=4)
keys = [M.fit_paramnames.index(k) for k in M.fitkeyparamnames]
B, cpr = 32, 512
n = B * cpr
d = M.d
# one long DE-MC burn-in gives both samplers the same (near-stationary) starting population
S = Sampler(M, n, kind="demc", seed=1, rungs=B, replicates=True)
S.adapt(True)
for _ in range(600):
    S.run(10); S.refresh_population()
S.adapt(False)
start = S.get()[0]
for kind, burn, rec, thin, kw in (("demc", 0, 4500, 15, {"single": False}), ("demc", 0, 4500, 15, {}), ("drj", 20, 100, 1, {})):
    T = Sampler(M, n, kind=kind, seed=7, rungs=B, replicates=True, theta0=start, **kw)
    if burn:
        T.adapt(True); T.run(burn); T.adapt(False)
    e0 = T.stats()["evals"]
    T.record(thin, rec // thin)
    t0 = time.perf_counter(); T.run(rec); dt = time.perf_counter() - t0
    st = T.stats()
    full, half = T.replicate_stats(keys), T.replicate_stats(keys, last=(rec // thin) // 2)
    ev = st["evals"] - e0
    ef, eh = np.median(full["ess"]), np.median(half["ess"])
    print(f"{kind} {kw}: {rec} iterations, {ev / n:.0f} evals per chain in {dt:.1f} s, acceptance {st['accept_rate']:.3f}; "
          f"ESS median whole window {ef:.0f}, first half {eh:.0f}; growth per 1000 evals per chain "
          f"{(ef - eh) / (ev / n / 2) * 1000:.0f}; split R-hat median {np.median(full['rhat']):.2f}", flush=True)
M.close()
