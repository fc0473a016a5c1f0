"""Load-balance diagnostics of the integrator launch at the bench size: per-mode step counts, isolated step latency, makespan bound."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks = np.linspace(0.01, 2000, 2019)
def run(ks, **kw):
    sb.solvept(prob, bg, ks, **kw); torch.cuda.synchronize(); t = time.time(); s = sb.solvept(prob, bg, ks, **kw); torch.cuda.synchronize(); return s, time.time() - t
s, t = run(ks)
att = s.stats[:, 0] + s.stats[:, 1]
print(f"2019 modes: {t*1e3:.1f} ms; attempts sum {att.sum()} max {att.max()} (k={ks[att.argmax()]:.0f}) mean {att.mean():.0f}; per-warp mean over 1184 warps {att.sum()/1184:.0f}")
for k in (2000.0, 1000.0, 100.0):
    s1, t1 = run(np.array([k]))
    a1 = s1.stats[0, 0] + s1.stats[0, 1]
    print(f"single mode k={k}: {t1*1e3:.1f} ms, {a1} attempts -> {t1/a1*1e6:.2f} us/step isolated")
for nct in (148 * 2, 148 * 4, 148 * 6, 148 * 8):
    s2, t2 = run(ks, nctas=nct)
    print(f"nctas={nct}: {t2*1e3:.1f} ms -> {t2/(att.sum()/nct)*1e6:.2f} us/step/warp if balanced")
