set -x
./profiles/microbench/lat_probe
python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k "drjacoby or replicate or k2_alone or age_model_stat or shipped" 2>&1 | tail -15
