python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "I290 or lane or golden" 2>&1 | grep -v "^$" | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
echo "== I290"; python bench.py --workload I290 --substeps 4 --theta narrow --steps 60 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
