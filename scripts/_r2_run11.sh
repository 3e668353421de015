set -x
python -m pytest tests/test_gpu_sampler.py -m gpu -q -s -k "replicate or k2_alone or age_model_stat or shipped" 2>&1 | grep -v "^$" | tail -40
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo rc=$?; tail -5 gpurun_out/r2_bench_n1_b.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_b.json 2>/dev/null; echo rc=$?
