# GPU run r2l: final build (slot-major basis sweep + reciprocal norm on) -- all GPU tests, smoke, bench, launch list, full ncu capture
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r2l.txt 2>&1; tail -2 gpurun_out/smoke_r2l.txt
(time python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2l.log 2>&1; tail -8 gpurun_out/gputest_r2l.log
python bench.py --steps 20 --warmup 5 --config4 512 > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; head -c 1000 gpurun_out/bench_r2l.json; tail -3 gpurun_out/bench_r2l.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2l.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r2l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 1 -o gpurun_out/prof_r2l python scripts/ab_fused.py --prof > gpurun_out/ncu_full_r2l.log 2>&1; tail -2 gpurun_out/ncu_full_r2l.log
python scripts/lanes_cost.py > gpurun_out/lanes_cost_r2l.txt 2>&1; cat gpurun_out/lanes_cost_r2l.txt
