# GPU run r2w: all GPU tests, smoke, bench (final build of the day: continuant pivots, branch-free factorisation, split kernel with controller warp)
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2w.log 2>&1; tail -4 gpurun_out/gputest_r2w.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --config4 0 > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2w.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['other_schedule'], d.get('config1_pk', {}).get('ms'), d.get('cl_default_chebyshev', {}).get('ms'))
PY
python - > gpurun_out/split_latency_r2w.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
f = lambda k: min(1e-2 / k, 1e-4)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ks100 = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
kc = sb.ChebyshevInterpolator(1e-2, 2e3, 60).xs
_, taus = sb.cmb_grids(bg)
print("split capacity", sb.split_capacity(prob))
for label, ks, kw in (("config 1: 100 log-spaced k, P(k) solve", ks100, dict(ptivini=f)), ("61 Chebyshev nodes, 300 save times, fused sources", kc, dict(saveat=taus, sources=dict(nS=2, scale_k=True), keep_states=False)),
                      ("252 strided modes of the 2019 (one rank's share at N = 8), fused sources", sb.cmb_grids(bg)[0][::8], dict(saveat=taus, sources=dict(nS=2, scale_k=True), keep_states=False))):
    a = timed(lambda: sb.solvept(prob, bg, ks, split=False, warn=False, **kw))
    b = timed(lambda: sb.solvept(prob, bg, ks, split=True, warn=False, **kw))
    s = sb.solvept(prob, bg, ks, split=True, warn=False, **kw)
    att = int((s.stats[:, 0] + s.stats[:, 1]).max())
    print(f"{label}: warp per mode {a:.2f} ms | CTA per mode {b:.2f} ms ({b / a:.2f}x); slowest mode {att} attempts -> {1e3 * a / att:.2f} / {1e3 * b / att:.2f} us per attempt", flush=True)
PY
cat gpurun_out/split_latency_r2w.txt
