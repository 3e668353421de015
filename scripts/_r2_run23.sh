EPI_ODE_LANES=2 python bench.py --chains 32768 --steps 10 --warmup 3 --no-sampler --no-cpu-baseline --no-secondary 2>&1 | tail -12 | cut -c1-400
EPI_ODE_LANES=2 python - <<'PY'
import numpy as np
from epi_mcmc_b200.model import Model, load_dataset
import os
M = Model(load_dataset("A300"), substeps=4)
mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
rng = np.random.default_rng(1)
for n in (24576, 28000, 32768, 40000, 65536):
    th = np.clip(init + (rng.uniform(size=(n, M.d)) - 0.5) * 0.01 * (mx - mn), mn, mx)
    try:
        a = M.calclogpost_batch(th); b = M.calclogpost_batch(th)
        print(n, "ok", np.array_equal(a[0], b[0], equal_nan=True), int((a[2] != 0).sum()), flush=True)
    except Exception as e:
        print(n, "ERR", e, flush=True)
PY
