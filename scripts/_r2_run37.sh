P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, "narrow", round(d["narrow_cloud"]["ms_per_step"],4))'
for n in 20480 24576 28160; do
  for l in 1 2; do
  echo "== chains $n lanes $l"; EPI_ODE_LANES=$l python bench.py --chains $n --steps 60 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
  done
done
B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 8 -o gpurun_out/r2_ncu_A300_8192 -f python bench.py $B --workload A300 --chains 8192 --substeps 4 --theta posterior > gpurun_out/ncu_A300_8192.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r2_ncu_A300_8192.ncu-rep > gpurun_out/r2_ncu_A300_8192_d.txt 2>/dev/null
python profiles/source_hotspots.py gpurun_out/r2_ncu_A300_8192.ncu-rep "k_ode_age2" "k_ode_age2ILi2ELb1E" 40 > gpurun_out/r2_d_k_ode_age2_pass2_lines.txt 2>gpurun_out/hs2.err
rm -f gpurun_out/r2_ncu_A300_8192.ncu-rep
