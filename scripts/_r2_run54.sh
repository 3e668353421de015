python bench.py --gpus 1 --steps 300 --warmup 5 > gpurun_out/r2_bench_n1_j.json 2> gpurun_out/r2_bench_n1_j.err; echo rc=$?; tail -2 gpurun_out/r2_bench_n1_j.err
