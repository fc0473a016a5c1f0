# GPU run r2p: split kernel after the restructuring (basis slots off warp 0's path, controller powers on two warps, corrected k_j private): bit-identity, cycle accounting, latency
mkdir -p gpurun_out
timeout 600 python scripts/split_diag.py > gpurun_out/split_diag_r2p.txt 2>&1; cat gpurun_out/split_diag_r2p.txt
timeout 300 python scripts/split_prof.py > gpurun_out/split_prof_r2p.txt 2>&1; cat gpurun_out/split_prof_r2p.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_kernel or fused or batched_cosmologies or retcodes" 2>&1 | tail -5
