# GPU run r2o: split kernel (one CTA of SB_R warps per mode for small launches): bit-identity test, all GPU tests, latency numbers
mkdir -p gpurun_out
timeout 600 python scripts/split_diag.py > gpurun_out/split_diag_r2o.txt 2>&1; cat gpurun_out/split_diag_r2o.txt
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/gputest_r2o.log 2>&1; tail -12 gpurun_out/gputest_r2o.log
python - > gpurun_out/split_latency_r2o.txt 2>&1 <<'PY'
import sys, time
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
cat gpurun_out/split_latency_r2o.txt
python bench.py --steps 5 --warmup 3 --config4 0 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; grep -o '"config1_pk.*' gpurun_out/bench_r2o.json | cut -c1-900; tail -3 gpurun_out/bench_r2o.err
