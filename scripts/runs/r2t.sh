# GPU run r2t: stage loop unrolled in the warp-per-mode kernel (default), optional in the split kernel: cycle accounting, all GPU tests, bench, split latency table
mkdir -p gpurun_out
timeout 300 python scripts/split_prof.py scripts/variants/prof.so > gpurun_out/split_prof_r2t.txt 2>&1; cat gpurun_out/split_prof_r2t.txt
timeout 300 python scripts/split_prof.py scripts/variants/prof_unroll.so > gpurun_out/split_prof_unroll_r2t.txt 2>&1; head -12 gpurun_out/split_prof_unroll_r2t.txt
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2t.log 2>&1; tail -5 gpurun_out/gputest_r2t.log
python bench.py --steps 10 --warmup 3 --config4 0 > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2t.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('config1_pk', {}).get('ms'), d.get('cl_default_chebyshev', {}).get('ms'))
PY
