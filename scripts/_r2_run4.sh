set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or well_cond or trajectories or edge" 2>&1 | tail -5
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()}, round(d["posterior_cloud"]["ms_per_step"],4), {k:round(v,4) for k,v in d["posterior_cloud"]["kernels"].items()})'
for n in 65536; do echo "== chains $n"; python bench.py --chains $n --steps 20 --warmup 3 --no-sampler --no-cpu-baseline 2>/dev/null | python -c "$P"; done
ncu --set full --clock-control none --import-source on -k regex:k_score -c 2 -o gpurun_out/r2_score_b -f python bench.py --chains 65536 --steps 1 --warmup 1 --no-sampler --no-cpu-baseline > gpurun_out/ncu_score_b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
