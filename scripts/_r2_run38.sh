python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane or both_ode or r_fallback or restart or window or tiny or golden or mirrored or substep" 2>&1 | grep -v "^$" | tail -4
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, "narrow", round(d["narrow_cloud"]["ms_per_step"],4))'
for n in 8192 32768 65536; do
  for mb in 4 5; do
  echo "== chains $n two lanes minb $mb"; EPI_ODE_LANES=2 EPI_ODE_MINB=$mb python bench.py --chains $n --steps 60 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
  done
done
