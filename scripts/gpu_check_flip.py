import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
from oracle import sbref
M = sb.ΛCDM(lmax=5)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
obg = sbref.Background.from_knots(sbref.planck18(lmax=5), bg.t, bg.y, bg.dy, bg.tau0, bg.kappa0)
ks = np.linspace(100, 2000, 96)
osol = sbref.solvept(obg, ks)
for nt in (2, 4, 8, 16, 64):
    sol = sb.solvept(prob, bg, ks, msub=nt)
    ds = sol.stats[:, 0] - osol["stats"][:, 0]
    rel = np.abs(sol.uend - osol["uend"]) / np.abs(osol["uend"]).max(axis=0, keepdims=True)
    print("msub", nt, "modes with different accepted-step count:", (ds != 0).sum(), "of", len(ks), "rejects differ:", (sol.stats[:, 1] != osol["stats"][:, 1]).sum(), "max rel-to-max uend diff %.2e median %.2e" % (rel.max(), np.median(rel.max(axis=1))))
