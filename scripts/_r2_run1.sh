set -x
python -m pytest tests -m gpu -x -q -k "lane or both_ode or r_fallback or golden" 2>&1 | tail -15
for n in 8192 16384 32768; do for l in 1 2 4; do
  echo "== chains $n lanes $l"; EPI_ODE_LANES=$l python bench.py --chains $n --steps 20 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels'], d['posterior_cloud']['ms_per_step'], d['posterior_cloud']['kernels'])"
done; done
