mkdir -p gpurun_out
timeout 900 python scripts/split_waves.py > gpurun_out/split_waves_r3v.txt 2>&1; cat gpurun_out/split_waves_r3v.txt | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_kernel or static_schedule" 2>&1 | tail -2
