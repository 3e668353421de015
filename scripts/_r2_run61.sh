python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -3
python bench.py --gpus 1 --steps 300 --warmup 5 > gpurun_out/r2_bench_n1_k.json 2> gpurun_out/r2_bench_n1_k.err; echo rc=$?; tail -2 gpurun_out/r2_bench_n1_k.err
