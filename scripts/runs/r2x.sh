# GPU run r2x: ncu evidence for the final kernels: launch list of a bench step, full capture of the scheduled bench launch, full captures of config 1 under both mappings
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2x.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r2x.log 2>&1; tail -2 gpurun_out/ncu_bench_r2x.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 1 -o gpurun_out/prof_r2x python scripts/ab_fused.py --prof > gpurun_out/ncu_full_r2x.log 2>&1; tail -2 gpurun_out/ncu_full_r2x.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 2 -o gpurun_out/prof_split_r2x python scripts/prof_split.py > gpurun_out/ncu_split_r2x.log 2>&1; tail -3 gpurun_out/ncu_split_r2x.log
ls -la gpurun_out/*.ncu-rep
