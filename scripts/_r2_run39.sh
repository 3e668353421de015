python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -4
python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/r2_bench_n1_h.json 2> gpurun_out/r2_bench_n1_h.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1_h.err
