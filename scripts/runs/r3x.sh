mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --config4 0 > gpurun_out/bench_r3x.json 2> gpurun_out/bench_r3x.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r3x.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, json.dumps(d['config1_pk_integrators'], indent=0)[:900])
PY
tail -2 gpurun_out/bench_r3x.err
