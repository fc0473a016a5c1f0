# GPU run r2q: continuant path pivots (both kernels), clears hoisted out of the split kernel's attempt: bit-identity, cycle accounting, all GPU tests, A/B of the warp-per-mode kernel
mkdir -p gpurun_out
timeout 600 python scripts/split_diag.py > gpurun_out/split_diag_r2q.txt 2>&1; tail -8 gpurun_out/split_diag_r2q.txt
timeout 300 python scripts/split_prof.py > gpurun_out/split_prof_r2q.txt 2>&1; cat gpurun_out/split_prof_r2q.txt
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2q.log 2>&1; tail -8 gpurun_out/gputest_r2q.log
python bench.py --steps 10 --warmup 3 --config4 0 > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2q.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('config1_pk', {}).get('ms'), d.get('cl_default_chebyshev', {}).get('ms'))
PY
cat gpurun_out/parity_r2.json | head -30
