# GPU run r3g: controller in log2/exp2 form + deferred basis (rows requested at the start of the attempt, swept before the first stage): A/B, bit-identity, GPU tests
mkdir -p gpurun_out
timeout 900 python scripts/ab2.py scripts/variants/before.so scripts/variants/defer.so > gpurun_out/ab_defer_r3g.txt 2>&1; tail -4 gpurun_out/ab_defer_r3g.txt
timeout 900 python scripts/ab_src.py scripts/variants/before.so scripts/variants/defer.so > gpurun_out/ab_src_defer_r3g.txt 2>&1; tail -4 gpurun_out/ab_src_defer_r3g.txt | cut -c1-330
timeout 300 python scripts/split_diag.py > gpurun_out/split_diag_r3g.txt 2>&1; grep -c "stats equal True, differing entries 0" gpurun_out/split_diag_r3g.txt; grep -c "usave equal True, S equal True" gpurun_out/split_diag_r3g.txt
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r3g.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r3g.log | tail -3; grep "^E \|^FAILED" gpurun_out/gputest_r3g.log | head -8
