B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
cap() { # name workload chains substeps theta
  ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 12 -o gpurun_out/r2_ncu_$1 -f \
    python bench.py $B --workload $2 --chains $3 --substeps $4 --theta $5 > gpurun_out/ncu_$1.log 2>&1
  python profiles/extract_ncu_constants.py gpurun_out/r2_ncu_$1.ncu-rep $2 $3 $4 $5 > /dev/null
  python profiles/summarize_ncu.py gpurun_out/r2_ncu_$1.ncu-rep > gpurun_out/r2_ncu_$1.txt 2>/dev/null
  rm -f gpurun_out/r2_ncu_$1.ncu-rep
}
cap X300_post X300 65536 4 posterior
cap A365_post A365 65536 4 posterior
cp profiles/ncu_constants.json gpurun_out/ncu_constants.json
