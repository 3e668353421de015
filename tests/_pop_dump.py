"""diagnostic (not collected): run the DE-MC sampler on a dataset (default A300) and dump its population (the theta
cloud a running sampler evaluates) + a Gaussian summary — the file bench.py reads as
epi_mcmc_b200/datasets/<workload>_posterior_gauss.json.  run: python tests/_pop_dump.py [iters] [dataset]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
name = sys.argv[2] if len(sys.argv) > 2 else "A300"
M = Model(load_dataset(name), substeps=4)
S = Sampler(M, 65536, kind="demc", seed=1)
S.run(iters)
theta, logl, logp = S.get()
os.makedirs("gpurun_out", exist_ok=True)
if name == "A300":
    np.save("gpurun_out/pop_A300.npy", theta.astype(np.float64))
mu, cov = theta.mean(0), np.cov(theta.T)
json.dump({"dataset": name, "iterations": iters, "chains": 65536, "mean": mu.tolist(), "cov": cov.tolist(),
           "logpost_mean": float((logl + logp).mean())}, open(f"gpurun_out/pop_{name}_gauss.json", "w"))
print("dumped", theta.shape, "logpost mean", (logl + logp).mean(), "accept", S.stats()["accept_rate"])
