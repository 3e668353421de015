python -m pytest tests/test_gpu_sampler.py -m gpu -q 2>&1 | grep -v "^$" | tail -6
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_g.json 2> gpurun_out/r2_bench_n1_g.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1_g.err
