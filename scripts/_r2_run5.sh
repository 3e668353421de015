set -x
./profiles/microbench/halfwarp
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, round(d["posterior_cloud"]["ms_per_step"],4))'
echo "== chains 65536"; python bench.py --chains 65536 --steps 20 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | python -c "$P"
for n in 8192 16384 32768; do for l in 1 4; do for w in 1 2 4 8; do
  echo "== chains $n lanes $l ways $w"; EPI_ODE_LANES=$l EPI_SPLIT_WAYS=$w python bench.py --chains $n --steps 40 --warmup 5 --no-sampler --no-cpu-baseline 2>/dev/null | python -c "$P"
done; done; done
