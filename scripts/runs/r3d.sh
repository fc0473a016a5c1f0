timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kencarp4" 2>&1 | grep -v "^$" | tail -40
