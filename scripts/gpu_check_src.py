import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
from oracle import sbref
M = sb.ΛCDM(lmax=5)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
obg = sbref.Background.from_knots(sbref.planck18(lmax=5), bg.t, bg.y, bg.dy, bg.tau0, bg.kappa0)
kint = sb.ChebyshevInterpolator(1e-2, 2e3, 60)
ks_fine, taus = sb.cmb_grids(bg)
ks = kint.xs[::6].copy()
S = sb.source_grid(prob, taus, ks, bg, scale_k=False)
us = S.sol.usave
oS = sbref.sources(obg, ks, taus, us)   # oracle sources from the GPU's saved states
Sg = S.dS.cpu().numpy()
for ik in range(len(ks)):
    a, b = Sg[ik, 0], oS[ik, :, 0]
    e = np.abs(a - b) / np.abs(b).max()
    a2, b2 = Sg[ik, 1, :-1], oS[ik, :-1, 1]
    e2 = np.abs(a2 - b2) / np.abs(b2).max()
    print("k=%8.2f ST err %.2e at it=%d (tau=%.4f)  SE err %.2e at it=%d" % (ks[ik], e.max(), e.argmax(), taus[e.argmax()], e2.max(), e2.argmax()))
# state comparison against the oracle's own solve, per variable
osol = sbref.solvept(obg, ks, saveat=taus)
d = np.abs(us - osol["usave"]); sc = np.abs(osol["usave"]).max(axis=1, keepdims=True) + 1e-300
print("per-variable max rel-to-max diff of saved states:", (d / sc).max(axis=(0, 1))[:20])
print("steps", S.sol.stats[:, 0], osol["stats"][:, 0])
