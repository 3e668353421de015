cat > /tmp/san_small.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from epi_mcmc_b200.model import Model, load_dataset
for name, n in (("A300", 300), ("I290", 200), ("S120", 500)):
    ds = load_dataset(name)
    M = Model(ds, substeps=4)
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    rng = np.random.default_rng(1)
    th = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.3 * (mx - mn), mn, mx)
    logl, logp, st = M.calclogpost_batch(th)
    print(name, np.isfinite(logl).mean(), flush=True)
    M.close()
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_small.py > gpurun_out/san_$tool.log 2>&1; echo "$tool rc=$?"; tail -6 gpurun_out/san_$tool.log
done
