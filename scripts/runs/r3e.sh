# GPU run r3e: all GPU tests, smoke, bench after the ESDIRK integrators and the non-inlined initial state; memcheck of the ESDIRK kernels
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r3e.log 2>&1; grep "passed\|failed" gpurun_out/gputest_r3e.log | tail -3; grep "^E " gpurun_out/gputest_r3e.log | head -5
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --config4 512 > gpurun_out/bench_r3e.json 2> gpurun_out/bench_r3e.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r3e.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['other_schedule']['ms_per_step'], d['config1_pk']['ms'], d['cl_default_chebyshev']['ms'], d['config4_sweep'].get('k_modes_per_s'), d['cpu_baseline']['value'])
PY
cat > /tmp/sdirk_small.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=5); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks = np.array([0.0, 0.5, 5.0, 40.0]); taus = np.geomspace(1e-3, bg.tau0, 9); taus[-1] = bg.t[-1]
for alg in ("TRBDF2", "KenCarp4"):
    s = sb.solvept(prob, bg, ks, alg=alg, maxiters=150, saveat=taus, sources=dict(nS=3, scale_k=True), warn=False); torch.cuda.synchronize()
    print(alg, s.retcode.tolist(), s.stats[:, 0].tolist())
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python /tmp/sdirk_small.py 2>&1 | tail -4
