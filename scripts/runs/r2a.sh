# GPU run r2a: all GPU tests (incl. the new bench-model parity tests and the fused-source bit-identity test), fused A/B, bench (both arms),
# launch list and one full ncu capture of the integrator with fused sources
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gputest_r2a.log 2>&1; tail -30 gpurun_out/gputest_r2a.log
python scripts/ab_fused.py 5 > gpurun_out/ab_fused_r2a.txt 2>&1; cat gpurun_out/ab_fused_r2a.txt
python bench.py --steps 5 --warmup 3 --config4 512 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 3500 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2a.json 2>&1; tail -c 1500 gpurun_out/bench_ref_r2a.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 1 -o gpurun_out/prof_r2a python scripts/ab_fused.py --prof > gpurun_out/ncu_full_r2a.log 2>&1; tail -3 gpurun_out/ncu_full_r2a.log
