python -m pytest tests -m gpu -x -q -k "lane or both_ode or r_fallback or golden or window or restart or two_round" 2>&1 | tail -5
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, round(d["posterior_cloud"]["ms_per_step"],4), {k:round(v,4) for k,v in d["posterior_cloud"]["kernels"].items()})'
for n in 8192 16384 32768; do for l in 1 4; do
  echo "== chains $n lanes $l"; EPI_ODE_LANES=$l python bench.py --chains $n --steps 20 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | python -c "$P"
done; done
for st in 2 4; do echo "== chains 65536 strips $st"; EPI_SCORE_STRIPS=$st python bench.py --chains 65536 --steps 20 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | python -c "$P"; done
