timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "alternative_integrators" 2>&1 | grep -v "^$" | tail -8 | cut -c1-300
