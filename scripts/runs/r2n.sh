# GPU run r2n: lanes test with the tight-tolerance derivative check; save-cost sweep of the schedule on the final build
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lockstep or restage or converged" -s > gpurun_out/gputest_r2n.log 2>&1; tail -15 gpurun_out/gputest_r2n.log
python scripts/ab_fused.py 5 > gpurun_out/ab_fused_r2n.txt 2>&1; cat gpurun_out/ab_fused_r2n.txt
