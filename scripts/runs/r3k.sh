# GPU run r3k: knot-time window in shared memory: A/B, cycle account, bit-identity, GPU tests
mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/before.so scripts/variants/tw.so > gpurun_out/ab_tw_r3k.txt 2>&1; tail -4 gpurun_out/ab_tw_r3k.txt
timeout 900 python scripts/ab_src.py scripts/variants/before.so scripts/variants/tw.so > gpurun_out/ab_src_tw_r3k.txt 2>&1; tail -2 gpurun_out/ab_src_tw_r3k.txt | cut -c1-330
timeout 600 python scripts/warp_prof.py > gpurun_out/warp_prof_r3k.txt 2>&1; head -14 gpurun_out/warp_prof_r3k.txt
timeout 300 python scripts/split_diag.py > gpurun_out/split_diag_r3k.txt 2>&1; grep -c "stats equal True, differing entries 0" gpurun_out/split_diag_r3k.txt; grep -c "usave equal True, S equal True" gpurun_out/split_diag_r3k.txt
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r3k.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r3k.log | tail -3; grep "^E \|^FAILED" gpurun_out/gputest_r3k.log | head -8
