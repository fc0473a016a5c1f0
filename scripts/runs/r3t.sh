timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "esdirk" 2>&1 | grep -v "^$" | tail -4 | cut -c1-300
