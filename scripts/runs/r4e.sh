timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "solution_object" 2>&1 | grep -v "^$" | tail -12 | cut -c1-300
