set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "X300" 2>&1 | grep -v "^$" | tail -40
