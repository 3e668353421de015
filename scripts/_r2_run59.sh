python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "X300 or waning or golden or other_substep or odd_period" 2>&1 | grep -v "^$" | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
for n in 8192 65536; do
  echo "== X300 posterior chains $n"; python bench.py --workload X300 --chains $n --substeps 4 --theta posterior --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
echo "== X300 narrow 65536"; python bench.py --workload X300 --substeps 4 --theta narrow --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
