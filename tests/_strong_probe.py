"""diagnostic (not collected): the strong-scaling shard (65,536 / N chains on one GPU) as free-running slices.
K disjoint chain ranges of the batch run on K streams with no join between consecutive steps: a range's step k+1 is
ordered behind its own step k only.  run: python tests/_strong_probe.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler

ds = load_dataset("A300")
g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "epi_mcmc_b200", "datasets", "A300_posterior_gauss.json")))
for lanes in ("0", "1"):
    os.environ["EPI_ODE_LANES"] = lanes
    M = Model(ds, substeps=4)
    mn, mx = M.df_params["min"], M.df_params["max"]
    for n in (8192, 16384, 32768, 65536):
        th = np.clip(np.random.default_rng(1).multivariate_normal(np.array(g["mean"]), np.array(g["cov"]), size=n), mn, mx)
        thd = torch.from_numpy(np.ascontiguousarray(th.T)).cuda()
        ll = torch.empty(n, dtype=torch.float64, device="cuda"); lp = torch.empty_like(ll)
        st = torch.empty(n, dtype=torch.int32, device="cuda")
        M.reserve(n)
        for ways in (1, 2, 4, 8):
            sl = ((n + ways - 1) // ways + 127) // 128 * 128
            streams = [torch.cuda.Stream() for _ in range(ways)]
            def step():
                for k, s in enumerate(streams):
                    f = k * sl; c = min(sl, n - f)
                    if c > 0:
                        M.eval_device_range(thd.data_ptr(), n, n, f, c, ll.data_ptr(), lp.data_ptr(), st.data_ptr(), s.cuda_stream)
            for _ in range(5): step()
            torch.cuda.synchronize()
            K = 200
            t0 = time.perf_counter()
            for _ in range(K): step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / K
            print(f"lanes {lanes} n {n:6d} ways {ways}: {dt*1e3:.4f} ms/step  {n/dt/1e6:.2f} M evals/s", flush=True)
    M.close()
os.environ["EPI_ODE_LANES"] = "0"
for ways in ("1", "2", "4", "8"):
    os.environ["EPI_SPLIT_WAYS"] = ways
    M = Model(ds, substeps=4)
    for n in (8192, 16384):
        S = Sampler(M, n, kind="demc", seed=1)
        S.run(200); S.stats()
        t0 = time.perf_counter(); S.run(500); S.stats(); dt = (time.perf_counter() - t0) / 500
        print(f"sampler ways {ways} n {n}: {dt*1e3:.4f} ms/iter {n/dt/1e6:.2f} M evals/s", flush=True)
    M.close()
