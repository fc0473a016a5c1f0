"""Profiling driver: one warm-up + one timed P(k) solve (ΛCDM lmax=10, 100 log-spaced k) — wrapped by ncu under gpurun."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 100
M = sb.ΛCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
ks = sb.loggrid(1e-4, 1.0, length=nk) / sb.k0 if nk <= 256 else np.linspace(0.01, 2000, nk)
for _ in range(2):
    P = sb.spectrum_matter(prob, ks, bgsol=bg)
torch.cuda.synchronize()
print("done", P[:3])
