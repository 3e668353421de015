P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
for n in 8192 16384 32768; do
for l in 1 2; do
  echo "== X300 posterior chains $n lanes $l"; EPI_ODE_LANES=$l python bench.py --workload X300 --chains $n --substeps 4 --theta posterior --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
done
