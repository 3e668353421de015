set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "quantile or marginalize" 2>&1 | grep -v "^$" | tail -8
python -m pytest tests/test_gpu_sampler.py -m gpu -q 2>&1 | grep -v "^$" | tail -8
