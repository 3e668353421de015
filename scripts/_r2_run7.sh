set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; echo rc=$?
python bench.py --chains 8192 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_8192_a.json 2>/dev/null; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_a_launches.csv python bench.py --steps 2 --warmup 1 --no-sampler --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 8 -o gpurun_out/r2_a_full -f python bench.py --chains 65536 --steps 1 --warmup 1 --no-sampler --no-cpu-baseline > gpurun_out/ncu_a_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
