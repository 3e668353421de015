set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, round(d["posterior_cloud"]["ms_per_step"],4), {k:round(v,4) for k,v in d["posterior_cloud"]["kernels"].items()})'
for n in 65536 8192; do echo "== chains $n"; python bench.py --chains $n --steps 20 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | python -c "$P"; done
for w in S365 I290 NE120; do echo "== $w"; python bench.py --workload $w --chains 65536 --steps 10 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), d["kernels"])'; done
