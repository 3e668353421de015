set -x
python -m pytest tests/test_gpu_sampler.py -m gpu -q -s -k "age_model_stat" 2>&1 | grep -v "^$" | tail -12
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mirrored" 2>&1 | grep -v "^$" | tail -25
bash scripts/_r2_ncu_all.sh
