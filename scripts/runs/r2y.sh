# GPU run r2y: all phase-0 paths on one half-warp (SB_PHALF): A/B against the previous build, bit-identity of the two mappings, GPU tests
mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/unroll0.so scripts/variants/phalf.so > gpurun_out/ab_phalf_r2y.txt 2>&1; tail -4 gpurun_out/ab_phalf_r2y.txt
timeout 900 python scripts/ab_src.py scripts/variants/unroll0.so scripts/variants/phalf.so > gpurun_out/ab_src_phalf_r2y.txt 2>&1; tail -4 gpurun_out/ab_src_phalf_r2y.txt
timeout 300 python scripts/split_diag.py > gpurun_out/split_diag_r2y.txt 2>&1; grep -c "stats equal True, differing entries 0" gpurun_out/split_diag_r2y.txt; grep "nS" gpurun_out/split_diag_r2y.txt
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2y.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r2y.log | tail -3
