# GPU run r3a: compute-sanitizer on the split kernel (memcheck, racecheck, synccheck), tiny bounded workload
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_split.py > gpurun_out/sanitize_split_plain_r3a.txt 2>&1; tail -6 gpurun_out/sanitize_split_plain_r3a.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_split.py > gpurun_out/sanitize_split_memcheck_r3a.txt 2>&1; tail -4 gpurun_out/sanitize_split_memcheck_r3a.txt
SB_SANITIZE_FAST=1 timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_split.py > gpurun_out/sanitize_split_synccheck_r3a.txt 2>&1; tail -4 gpurun_out/sanitize_split_synccheck_r3a.txt
SB_SANITIZE_FAST=1 timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python scripts/sanitize_split.py > gpurun_out/sanitize_split_racecheck_r3a.txt 2>&1; grep -c "hazard" gpurun_out/sanitize_split_racecheck_r3a.txt; tail -30 gpurun_out/sanitize_split_racecheck_r3a.txt | cut -c1-250
