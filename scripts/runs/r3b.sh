mkdir -p gpurun_out
timeout 600 python scripts/split_diag.py > gpurun_out/split_diag_r3b.txt 2>&1; grep -c "differing entries 0 of" gpurun_out/split_diag_r3b.txt; grep "nS\|nx=8" gpurun_out/split_diag_r3b.txt | cut -c1-200
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r3b.log 2>&1; grep "passed\|failed\|Error" gpurun_out/gputest_r3b.log | tail -5
