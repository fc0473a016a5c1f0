# GPU run r2c: all GPU tests (lockstep lanes new), cost of the lanes, sanitizer on the small workload, bench
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2c.log 2>&1; tail -40 gpurun_out/gputest_r2c.log
python scripts/lanes_cost.py > gpurun_out/lanes_cost_r2c.txt 2>&1; cat gpurun_out/lanes_cost_r2c.txt
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/memcheck_r2c.txt 2>&1; tail -6 gpurun_out/memcheck_r2c.txt
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/racecheck_r2c.txt 2>&1; grep -c "hazard" gpurun_out/racecheck_r2c.txt; tail -4 gpurun_out/racecheck_r2c.txt
python bench.py --steps 5 --warmup 3 --config4 512 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 1500 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
