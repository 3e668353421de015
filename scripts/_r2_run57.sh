python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -6
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
for l in 1 2; do
  echo "== X300 narrow lanes $l"; EPI_ODE_LANES=$l python bench.py --workload X300 --substeps 4 --theta narrow --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
cp gpurun_out/pop_X300_gauss.json epi_mcmc_b200/datasets/X300_posterior_gauss.json 2>/dev/null || python tests/_pop_dump.py 4000 X300 | tail -1
cp gpurun_out/pop_X300_gauss.json epi_mcmc_b200/datasets/X300_posterior_gauss.json
for l in 1 2; do
  echo "== X300 posterior lanes $l"; EPI_ODE_LANES=$l python bench.py --workload X300 --substeps 4 --theta posterior --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
