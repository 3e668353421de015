python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "piece_sequence or mirrored or lane_mappings or restart or window or golden or tiny" 2>&1 | grep -v "^$" | tail -6
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, "narrow", round(d["narrow_cloud"]["ms_per_step"],4))'
for f in 0 1; do
  echo "== EPI_ODE_FLOW=$f"; EPI_ODE_FLOW=$f python bench.py --steps 100 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
for n in 32768 49152; do
  for f in 0 1; do
  echo "== chains $n EPI_ODE_FLOW=$f"; EPI_ODE_LANES=1 EPI_ODE_FLOW=$f python bench.py --chains $n --steps 100 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
  done
done
