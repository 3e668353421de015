python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_stream or lane or both_ode or golden or edge" 2>&1 | grep -v "^$" | tail -6
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, "narrow", round(d["narrow_cloud"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e6,2))'
for n in 4096 8192 16384 24576; do for ts in 1 0; do
  echo "== chains $n two_stream $ts"; EPI_TWO_STREAM=$ts python bench.py --chains $n --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done; done
