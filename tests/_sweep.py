"""one-off wide parity sweep (not collected by pytest): GPU vs oracle on box-uniform + near-init proposals"""
import sys, time, json, os
SEED = int(os.environ.get('SWEEP_SEED', '0'))
BAD = []
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from epi_mcmc_b200.model import Model, load_dataset
from oracle import oracle
from conftest import logl_within_bar, rel_err
for name, n, m in (("A300", 24576, 4), ("A300", 8192, 1), ("I290", 16384, 4), ("A365", 8192, 2), ("S120", 32768, 4),
                   ("NE120", 32768, 3), ("S365", 16384, 4), ("DE", 16384, 4), ("SE_NE", 16384, 4)):
    ds = load_dataset(name)
    M = Model(ds, substeps=m)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(sum(map(ord, name)) * 7 + m + SEED)
    box = mn + rng.uniform(size=(n // 2, M.d)) * (mx - mn)
    near = np.clip(init + (rng.uniform(size=(n - n // 2, M.d)) - 0.5) * 0.4 * (mx - mn), mn, mx)
    theta = np.vstack([box, near])
    t0 = time.time()
    logl, logp, status = M.calclogpost_batch(theta)
    n_fallback = M.fallback_count()
    off, pad, terms = M.eval_detail(theta)
    ref = oracle.evaluate(ds["flavour"], ds, theta, ds["period"], m, n_threads=16)
    cond = oracle.conditioning(ds["flavour"], ds, theta, ds["period"], m, ref["logl"], n_threads=16)
    ok_int = np.array_equal(status, ref["status"]) and np.array_equal(off, ref["offset"]) and np.array_equal(pad, ref["padding"])
    within = logl_within_bar(logl, ref["logl"], cond, 1e-10)
    strict = rel_err(logl, ref["logl"]) <= 1e-10
    lp_ok = rel_err(logp, ref["logp"]).max()
    nanmis = int((np.isnan(logl) != np.isnan(ref["logl"])).sum())
    print(f"{name} m={m} n={n}: ints_equal={ok_int} within_bar={within.mean():.5f} strict_1e-10={np.mean(strict):.4f} "
          f"logp_max_rel={lp_ok:.2e} nan_mismatch={nanmis} status0={np.mean(status==0):.3f} r_fallback={n_fallback} t={time.time()-t0:.1f}s", flush=True)
    if not ok_int:
        bad = np.nonzero((status != ref["status"]) | (off != ref["offset"]) | (pad != ref["padding"]))[0]
        print("  int mismatches:", len(bad), [(int(i), int(status[i]), int(ref["status"][i]), int(off[i]), int(ref["offset"][i])) for i in bad[:5]])
        for i in bad[:20]:
            BAD.append(dict(name=name, m=m, theta=theta[i].tolist(), gpu_status=int(status[i]), ref_status=int(ref["status"][i]),
                            gpu_logl=None if np.isnan(logl[i]) else float(logl[i]), ref_logl=None if np.isnan(ref["logl"][i]) else float(ref["logl"][i]),
                            gpu_terms=[None if np.isnan(v) else float(v) for v in terms[i][:6]], ref_terms=[None if np.isnan(v) else float(v) for v in ref["terms"][i][:6]],
                            off=int(off[i]), pad=int(pad[i])))
    M.close()

json.dump(BAD, open('/root/repo/gpurun_out/sweep_bad.json', 'w'))
