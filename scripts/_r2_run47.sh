python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
echo "== NE120 4096"; python bench.py --workload NE120 --chains 4096 --theta narrow --steps 200 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
echo "== S120 1M"; python bench.py --workload S120 --chains 1048576 --theta narrow --steps 20 --warmup 3 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
echo "== S365 1M"; python bench.py --workload S365 --chains 1048576 --theta narrow --steps 10 --warmup 3 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
