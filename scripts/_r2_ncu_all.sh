# ncu --set full captures of bench.py's own workloads (first launch of every kernel role) + K2 at 2^22 chains
set -x
B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
cap() { # name workload chains substeps theta
  ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 8 -o gpurun_out/r2_ncu_$1 -f \
    python bench.py $B --workload $2 --chains $3 --substeps $4 --theta $5 > gpurun_out/ncu_$1.log 2>&1
}
cap A300_post A300 65536 4 posterior
cap A300_m1 A300 65536 1 posterior
cap A365 A365 65536 4 narrow
cap I290 I290 65536 4 narrow
cap S365 S365 1048576 4 narrow
cap S120 S120 1048576 4 narrow
cap NE120 NE120 4096 4 narrow
ncu --set full --clock-control none --import-source on -k regex:"k_accept_propose" -c 3 -o gpurun_out/r2_ncu_k2 -f \
  python -c "
import ctypes as C
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200 import _capi
M = Model(load_dataset('A300'), substeps=4)
us, by, ar = C.c_double(), C.c_double(), C.c_double()
_capi.check(M._lib.epi_k2_bench(M._h, 1, 1 << 22, 2, C.byref(us), C.byref(by), C.byref(ar)), M._h)
" > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out/r2_ncu_*.ncu-rep
