python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane or both_ode or r_fallback or restart or window or tiny or golden or mirrored or substep" 2>&1 | grep -v "^$" | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
for cfg in "A300 1 posterior" "I290 4 narrow" "A365 4 narrow"; do
  set -- $cfg
  for l in 1 2; do
    echo "== $1 m=$2 lanes $l"; EPI_ODE_LANES=$l python bench.py --workload $1 --substeps $2 --theta $3 --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
  done
done
B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
cap() { # name workload chains substeps theta keep
  ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 12 -o gpurun_out/r2_ncu_$1 -f \
    python bench.py $B --workload $2 --chains $3 --substeps $4 --theta $5 > gpurun_out/ncu_$1.log 2>&1
  python profiles/extract_ncu_constants.py gpurun_out/r2_ncu_$1.ncu-rep $2 $3 $4 $5 > /dev/null
  python profiles/summarize_ncu.py gpurun_out/r2_ncu_$1.ncu-rep > gpurun_out/r2_ncu_$1.txt 2>/dev/null
  if [ "$6" != keep ]; then rm -f gpurun_out/r2_ncu_$1.ncu-rep; fi
}
cap A300_post A300 65536 4 posterior keep
python profiles/source_hotspots.py gpurun_out/r2_ncu_A300_post.ncu-rep "k_ode_age2" "k_ode_age2ILi2ELb1E" 40 > gpurun_out/r2_e_k_ode_age2_pass2_lines.txt 2>/dev/null
rm -f gpurun_out/r2_ncu_A300_post.ncu-rep
cap A300_m1 A300 65536 1 posterior
cap A365 A365 65536 4 narrow
cap I290 I290 65536 4 narrow
cap A300_narrow A300 65536 4 narrow
cp profiles/ncu_constants.json gpurun_out/ncu_constants.json
