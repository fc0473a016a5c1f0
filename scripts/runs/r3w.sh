mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r3w.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r3w.log | tail -3; grep "^E \|^FAILED" gpurun_out/gputest_r3w.log | head -8
python bench.py --steps 10 --warmup 3 --config4 0 > gpurun_out/bench_r3w.json 2> gpurun_out/bench_r3w.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r3w.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['config1_pk']['ms'], d['cl_default_chebyshev']['ms'])
PY
