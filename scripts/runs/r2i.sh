# GPU run r2i: config-4 failure-set parity against the oracle, bench at the driver's N = 1 command line (full config-4 sweep)
mkdir -p gpurun_out
python scripts/sweep_fail_oracle.py 512 > gpurun_out/config4_failures_r2i.txt 2>&1; tail -12 gpurun_out/config4_failures_r2i.txt
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; head -c 600 gpurun_out/bench_r2i.json; grep -o '"config4_sweep.*' gpurun_out/bench_r2i.json | cut -c1-700; tail -4 gpurun_out/bench_r2i.err
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/bench_ref_r2i.json 2> gpurun_out/bench_ref_r2i.err; tail -c 400 gpurun_out/bench_ref_r2i.json; tail -4 gpurun_out/bench_ref_r2i.err
