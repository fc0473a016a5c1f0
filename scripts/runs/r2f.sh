# GPU run r2f: all GPU tests (closing step of the lanes, solution object, communicator, nx=16), bench with the TMA + reciprocal defaults,
# launch list + full ncu capture of the integrator, lanes cost
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2f.log 2>&1; tail -25 gpurun_out/gputest_r2f.log
python bench.py --steps 5 --warmup 3 --config4 512 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -c 2500 gpurun_out/bench_r2f.json; tail -5 gpurun_out/bench_r2f.err
python scripts/lanes_cost.py > gpurun_out/lanes_cost_r2f.txt 2>&1; cat gpurun_out/lanes_cost_r2f.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r2f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 1 -o gpurun_out/prof_r2f python scripts/ab_fused.py --prof > gpurun_out/ncu_full_r2f.log 2>&1; tail -3 gpurun_out/ncu_full_r2f.log
