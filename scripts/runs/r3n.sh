# GPU run r3n: reciprocal-based slot weights: A/B, bit-identity, GPU tests
mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/before.so scripts/variants/nodiv.so > gpurun_out/ab_nodiv_r3n.txt 2>&1; tail -4 gpurun_out/ab_nodiv_r3n.txt
timeout 900 python scripts/ab_src.py scripts/variants/before.so scripts/variants/nodiv.so > gpurun_out/ab_src_nodiv_r3n.txt 2>&1; tail -2 gpurun_out/ab_src_nodiv_r3n.txt | cut -c1-330
timeout 300 python scripts/split_diag.py > gpurun_out/split_diag_r3n.txt 2>&1; grep -c "stats equal True, differing entries 0" gpurun_out/split_diag_r3n.txt; grep -c "usave equal True, S equal True" gpurun_out/split_diag_r3n.txt
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r3n.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r3n.log | tail -3; grep "^E \|^FAILED" gpurun_out/gputest_r3n.log | head -8
