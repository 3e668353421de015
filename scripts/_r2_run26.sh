set -x
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -5
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_f.json 2> gpurun_out/r2_bench_n1_f.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1_f.err
