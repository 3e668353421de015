python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane or both_ode or r_fallback or restart or window or two_round or tiny" 2>&1 | grep -v "^$" | tail -6
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, "narrow", round(d["narrow_cloud"]["ms_per_step"],4))'
for n in 4096 8192 12288 16384 24576 32768; do for l in 1 2 4; do
  echo "== chains $n lanes $l"; EPI_ODE_LANES=$l python bench.py --chains $n --steps 40 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done; done
