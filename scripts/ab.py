"""A/B timing of integrator builds on the SAME box: python scripts/ab.py scripts/variants/a.so scripts/variants/b.so ..."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks100 = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
ks2k = np.linspace(0.01, 2000, 2019)
f = lambda k: min(1e-2 / k, 1e-4)
libs = [(p, C.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
ref = None
for rnd in range(3):
    for p, lib in libs:
        prob.lib = lib
        def run(ks, ptiv):
            torch.cuda.synchronize(); t = time.time(); s = sb.solvept(prob, bg, ks, ptivini=ptiv); torch.cuda.synchronize(); return s, time.time() - t
        s1, t1 = run(ks100, f); s2, t2 = run(ks2k, -np.inf)
        if ref is None: ref = s2.uend.copy()
        dev = np.nanmax(np.abs(s2.uend - ref)) / np.abs(ref).max()
        print(f"round {rnd} {os.path.basename(p):28s} P(k) 100 modes {t1*1e3:7.1f} ms | 2019 modes {t2*1e3:7.1f} ms | steps {s2.stats[:,0].sum()} ok {s2.success} dev-vs-first {dev:.1e}")
