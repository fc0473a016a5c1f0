# GPU run r3p: final verification of the round (all GPU tests, smoke, bench 20 steps with the config-4 sample, launch list)
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r3p.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r3p.log | tail -3; grep "^E \|^FAILED" gpurun_out/gputest_r3p.log | head -8
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --config4 512 > gpurun_out/bench_r3p.json 2> gpurun_out/bench_r3p.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r3p.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['other_schedule']['ms_per_step'], d['config1_pk']['ms'], d['cl_default_chebyshev']['ms'], d['config4_sweep'].get('k_modes_per_s'), d['cpu_baseline']['value'], d['clocks'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r3p.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r3p.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sb_integrate -s 2 -c 1 -o gpurun_out/prof_r3p python scripts/ab_fused.py --prof > gpurun_out/ncu_full_r3p.log 2>&1; tail -1 gpurun_out/ncu_full_r3p.log
timeout 600 python scripts/lanes_cost.py > gpurun_out/lanes_cost_r3p.txt 2>&1; tail -6 gpurun_out/lanes_cost_r3p.txt | cut -c1-250
