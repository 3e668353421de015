# ncu --set full captures of bench.py's own workloads (first launch of every kernel role) + K2 at 2^22 chains.
# Only the text summaries and profiles/ncu_constants.json travel back (gpurun_out/ is capped at 64 MiB): the reports
# are summarised on the box and deleted, except the headline workload's and K2's.
B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
cap() { # name workload chains substeps theta keep
  ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 12 -o gpurun_out/r2_ncu_$1 -f \
    python bench.py $B --workload $2 --chains $3 --substeps $4 --theta $5 > gpurun_out/ncu_$1.log 2>&1
  python profiles/extract_ncu_constants.py gpurun_out/r2_ncu_$1.ncu-rep $2 $3 $4 $5 > /dev/null
  python profiles/summarize_ncu.py gpurun_out/r2_ncu_$1.ncu-rep > gpurun_out/r2_ncu_$1.txt 2>/dev/null
  if [ "$6" != keep ]; then rm -f gpurun_out/r2_ncu_$1.ncu-rep; fi
}
cap A300_post A300 65536 4 posterior keep
python profiles/source_hotspots.py gpurun_out/r2_ncu_A300_post.ncu-rep "k_score" "k_scoreILi2E" 45 > gpurun_out/r2_b_k_score_lines.txt 2>/dev/null
python profiles/source_hotspots.py gpurun_out/r2_ncu_A300_post.ncu-rep "k_ode" "k_odeILi2ELb1ELb0ELb0E" 40 > gpurun_out/r2_b_k_ode_pass2_lines.txt 2>/dev/null
python profiles/source_hotspots.py gpurun_out/r2_ncu_A300_post.ncu-rep "k_mid" "k_midILi2E" 40 > gpurun_out/r2_b_k_mid_lines.txt 2>/dev/null
rm -f gpurun_out/r2_ncu_A300_post.ncu-rep
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
python profiles/summarize_ncu.py gpurun_out/r2_ncu_k2.ncu-rep > gpurun_out/r2_ncu_k2.txt 2>/dev/null
cp profiles/ncu_constants.json gpurun_out/ncu_constants.json
du -sh gpurun_out
