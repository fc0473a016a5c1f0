# GPU run r2g: A/B of the half-warp-aware path lanes (generator change only), all GPU tests, bench, full ncu capture
mkdir -p gpurun_out
python scripts/ab2.py scripts/variants/r2g_oldlanes.so scripts/variants/r2g_halfwarp.so > gpurun_out/ab_lanes_r2g.txt 2>&1; cat gpurun_out/ab_lanes_r2g.txt
(time python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2g.log 2>&1; tail -12 gpurun_out/gputest_r2g.log
python bench.py --steps 5 --warmup 3 --config4 512 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; head -c 1200 gpurun_out/bench_r2g.json; tail -5 gpurun_out/bench_r2g.err
ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 1 -o gpurun_out/prof_r2g python scripts/ab_fused.py --prof > gpurun_out/ncu_full_r2g.log 2>&1; tail -3 gpurun_out/ncu_full_r2g.log
