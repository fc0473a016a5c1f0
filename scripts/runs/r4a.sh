# GPU run r4a: compute-sanitizer memcheck on the last build: small workload over all kernels (fast mode), split kernel, ESDIRK kernels
mkdir -p gpurun_out
SB_SANITIZE_FAST=1 timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck_r4a.txt 2>&1; tail -4 gpurun_out/sanitize_memcheck_r4a.txt | cut -c1-200
SB_SANITIZE_FAST=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_split.py > gpurun_out/sanitize_split_memcheck_r4a.txt 2>&1; tail -3 gpurun_out/sanitize_split_memcheck_r4a.txt | cut -c1-200
