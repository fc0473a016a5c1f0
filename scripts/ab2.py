"""A/B timing of integrator builds on the SAME box, latency regime (100 modes), bench size (2019 modes) and throughput regime (3 x 2019 modes):
python scripts/ab2.py scripts/variants/a.so scripts/variants/b.so ..."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks100 = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
ks2k = np.linspace(0.01, 2000, 2019)
ks6k = np.linspace(0.01, 2000, 3 * 2019)
f = lambda k: min(1e-2 / k, 1e-4)
libs = [(p, C.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
ref = None
for p, lib in libs:
    lib.sbm_resident_warps.restype = C.c_int
for rnd in range(2):
    for p, lib in libs:
        prob.lib = lib
        def run(ks, ptiv):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); s = sb.solvept(prob, bg, ks, ptivini=ptiv); b.record(); torch.cuda.synchronize(); return s, a.elapsed_time(b)
        s1, t1 = run(ks100, f); s2, t2 = run(ks2k, -np.inf); s3, t3 = run(ks6k, -np.inf)
        if ref is None: ref = s2.uend.copy()
        dev = np.nanmax(np.abs(s2.uend - ref)) / np.abs(ref).max()
        att = (s3.stats[:, 0] + s3.stats[:, 1]).sum()
        print(f"round {rnd} {os.path.basename(p):20s} warps {lib.sbm_resident_warps():5d} | 100 modes {t1:7.1f} ms | 2019 modes {t2:7.1f} ms | 6057 modes {t3:7.1f} ms = {att/t3/1e3:6.2f} M attempts/s | ok {s2.success and s3.success} dev-vs-first {dev:.1e}")
