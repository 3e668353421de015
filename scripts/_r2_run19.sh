set -x
python -m pytest tests/test_gpu_sampler.py -m gpu -q -s -k "age_model_stat" 2>&1 | grep -v "^$" | tail -8
python tests/_ess_compare.py 2>&1 | tail -4
