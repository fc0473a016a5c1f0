"""Where does the split mapping stop paying?  Launches with more modes than CTAs fit (second wave from the queue) against one warp per mode.  GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
kf, taus = sb.cmb_grids(bg)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
kw = dict(saveat=taus, sources=dict(nS=2, scale_k=True), keep_states=False, warn=False)
print("capacity", sb.split_capacity(prob))
for stride in (8, 6, 5, 4, 3):
    ks = kf[::stride]
    w = timed(lambda: sb.solvept(prob, bg, ks, split=False, **kw)); s = timed(lambda: sb.solvept(prob, bg, ks, split=True, **kw))
    print(f"C_l grid, every {stride}th mode: {len(ks):4d} modes: warp per mode {w:6.2f} ms | CTA per mode {s:6.2f} ms", flush=True)
f = lambda k: min(1e-2 / k, 1e-4)
for n in (300, 450, 592, 800):
    ks = sb.loggrid(1e-4, 1.0, length=n) / sb.k0
    w = timed(lambda: sb.solvept(prob, bg, ks, ptivini=f, split=False, warn=False)); s = timed(lambda: sb.solvept(prob, bg, ks, ptivini=f, split=True, warn=False))
    print(f"P(k) grid, {n:4d} log-spaced modes: warp per mode {w:6.2f} ms | CTA per mode {s:6.2f} ms", flush=True)
