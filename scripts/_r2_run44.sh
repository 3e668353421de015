python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -4
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, "narrow", round(d["narrow_cloud"]["ms_per_step"],4))'
for n in 8192 65536; do
  echo "== chains $n"; python bench.py --chains $n --steps 60 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
