"""diagnostic (not collected): run the DE-MC sampler on A300 and dump its population (the theta cloud a running
sampler evaluates) + a Gaussian summary.  run: python tests/_pop_dump.py [iters]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
M = Model(load_dataset("A300"), substeps=4)
S = Sampler(M, 65536, kind="demc", seed=1)
S.run(iters)
theta, logl, logp = S.get()
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/pop_A300.npy", theta.astype(np.float64))
mu, cov = theta.mean(0), np.cov(theta.T)
json.dump({"dataset": "A300", "iterations": iters, "chains": 65536, "mean": mu.tolist(), "cov": cov.tolist(),
           "logpost_mean": float((logl + logp).mean())}, open("gpurun_out/pop_A300_gauss.json", "w"))
print("dumped", theta.shape, "logpost mean", (logl + logp).mean(), "accept", S.stats()["accept_rate"])
