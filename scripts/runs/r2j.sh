# GPU run r2j: verification of the final tree -- smoke, all GPU tests, memcheck on the small workload, failure-set check, bench
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r2j.txt 2>&1; tail -3 gpurun_out/smoke_r2j.txt
(time python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2j.log 2>&1; tail -12 gpurun_out/gputest_r2j.log
SB_SANITIZE_FAST=1 timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/memcheck_r2j.txt 2>&1; tail -4 gpurun_out/memcheck_r2j.txt
python scripts/sweep_fail_oracle.py 512 > gpurun_out/config4_failures_r2j.txt 2>&1; tail -8 gpurun_out/config4_failures_r2j.txt
python bench.py --steps 5 --warmup 3 --config4 512 > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; head -c 700 gpurun_out/bench_r2j.json; tail -3 gpurun_out/bench_r2j.err
