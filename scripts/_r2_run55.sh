for ds in I290 A365 X300; do python tests/_pop_dump.py 4000 $ds 2>&1 | tail -1; done
cp gpurun_out/pop_I290_gauss.json epi_mcmc_b200/datasets/I290_posterior_gauss.json
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
for l in 1 2; do
  echo "== I290 posterior lanes $l"; EPI_ODE_LANES=$l python bench.py --workload I290 --substeps 4 --theta posterior --steps 60 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
echo "== I290 posterior default"; python bench.py --workload I290 --substeps 4 --theta posterior --steps 60 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
