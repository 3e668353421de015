"""diagnostic (not collected): where a sampler iteration's time goes — eval on the sampler's own population vs the
iteration as a whole.  run: python tests/_samp_overhead.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler

M = Model(load_dataset("A300"), substeps=4)
n = 65536
S = Sampler(M, n, kind="demc", seed=1)
for it in (0, 1000, 2000):
    if it:
        S.run(1000)
    S.get()
    t0 = time.perf_counter(); S.run(100); S.stats(); S.get(); dt = (time.perf_counter() - t0) / 100
    theta = S.get()[0]
    M.set_timing(True)
    for _ in range(3):
        l, p, s = M.calclogpost_batch(theta)
    k = M.last_kernel_ms()
    M.set_timing(False)
    print(f"after {it + 100:5d} iterations: {dt * 1e3:.3f} ms/iteration; eval kernels on the current population {np.round(k, 3)} sum {sum(k):.3f} ms; "
          f"status!=0 {int((s != 0).sum())}, fallbacks {M.fallback_count()}", flush=True)

# where does the population's extra pass-2 time come from?  same chains, different arrangement
theta = S.get()[0]
off, pad, _ = M.eval_detail(theta)
def timed(th, label):
    M.set_timing(True)
    for _ in range(3):
        M.calclogpost_batch(th)
    k = M.last_kernel_ms(); M.set_timing(False)
    print(f"{label:46s} {np.round(k, 3)}", flush=True)
timed(theta, "population as is")
timed(theta[np.argsort(off, kind='stable')], "sorted by data_offset")
key = off * 1000 + np.floor(theta[:, 21] * 4)  # a second breakpoint parameter
timed(theta[np.lexsort((theta[:, 22], theta[:, 21], off))], "sorted by offset, theta21, theta22")
timed(np.repeat(theta[:2048], 32, axis=0), "every warp = 32 copies of one chain")
timed(np.tile(theta[:1], (n, 1)), "all chains identical")
print("offset spread", np.percentile(off, [1, 25, 50, 75, 99]), "padding", np.percentile(pad, [1, 50, 99]))
# how many days does a warp see with a breakpoint in at least one lane?  (2q schedule, model-age-groups2q.R:184-197)
if M.flavour == "age2q":
    ds = M.data
    lo_, ltp = ds["lockdown_offset"], ds["lockdown_transition_period"]
    o = off.astype(float)
    t2 = o + lo_ + ltp; t3 = t2 + theta[:, 23]; t4 = t3 + theta[:, 24]; t5 = o + ds["d5"]; t6 = t5 + theta[:, 31]
    t7 = t6 + theta[:, 35]; t8 = o + ds["d8"]
    bp = np.stack([o + lo_ + theta[:, 18], o + lo_ + np.maximum(theta[:, 18], 0), t2, t3, t4, t5, t6, t7, t8, t8 + 7], 1)
    days = np.floor(bp).astype(int).reshape(-1, 32 * 10)
    print("distinct cut days per warp: mean", np.mean([len(np.unique(r)) for r in days[:512]]),
          " per-breakpoint sd (days):", np.round(bp.std(0), 2))
