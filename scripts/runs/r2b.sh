# GPU run r2b: GPU tests, fused A/B with the save-cost sweep, bench (both arms), launch list, full ncu capture of the fused integrator
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2b.log 2>&1; tail -40 gpurun_out/gputest_r2b.log
python scripts/ab_fused.py 5 > gpurun_out/ab_fused_r2b.txt 2>&1; cat gpurun_out/ab_fused_r2b.txt
python bench.py --steps 5 --warmup 3 --config4 512 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 3500 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 1 -o gpurun_out/prof_r2b python scripts/ab_fused.py --prof > gpurun_out/ncu_full_r2b.log 2>&1; tail -3 gpurun_out/ncu_full_r2b.log
