# GPU run r4d: last full check of the committed state: all GPU tests, smoke, default bench command line
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -x -q -m gpu) > gpurun_out/gputest_r4d.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r4d.log | tail -3; grep "^E \|^FAILED" gpurun_out/gputest_r4d.log | head -8
python __graft_entry__.py smoke 2>&1 | tail -1
(time python bench.py) > gpurun_out/bench_r4d.json 2> gpurun_out/bench_r4d.err; tail -4 gpurun_out/bench_r4d.err | cut -c1-200; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r4d.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches')}, d['e2e'], d['roofline']['frac'], d['config4_sweep'].get('k_modes_per_s'), sorted(d.keys()))
PY
