"""A/B of integrator builds on the bench workload proper: 2019 modes, 300 save times, sources formed at the save times (queue order and,
with a cost model learnt from the first solve, the static schedule).  python scripts/ab_src.py a.so b.so ...   GPU box only."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks, taus = sb.cmb_grids(bg)
libs = [(p, C.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); r = fn(); b.record(); torch.cuda.synchronize(); return r, a.elapsed_time(b)
ref = None
for rnd in range(2):
    for p, lib in libs:
        prob.lib = lib
        kw = dict(saveat=taus, sources=dict(nS=2, scale_k=True), keep_states=False, warn=False)
        s0, t0 = timed(lambda: sb.solvept(prob, bg, ks, warn=False))
        s1, t1 = timed(lambda: sb.solvept(prob, bg, ks, **kw))
        att = (s1.stats[:, 0] + s1.stats[:, 1]).astype(np.float64)
        s3, t3 = timed(lambda: sb.solvept(prob, bg, ks, cost=att, warn=False))
        extra = ""
        for sc in (0.2, 0.35, 0.5, 0.7, 1.0):
            s2, t2 = timed(lambda: sb.solvept(prob, bg, ks, cost=att + sc * len(taus), **kw))
            extra += f" {sc}: {t2:.1f}"
        extra += f" | static schedule without save times {t3:.1f} ms"
        S = s2.d_S.cpu().numpy()
        if ref is None: ref = S.copy()
        print(f"round {rnd} {os.path.basename(p):14s} | no save times {t0:6.1f} ms | 300 save times + sources: queue {t1:6.1f} ms, static schedule by save_cost{extra} | ok {s1.success and s2.success} "
              f"dev-vs-first {np.nanmax(np.abs(S - ref)) / np.nanmax(np.abs(ref)):.1e}", flush=True)
