import sys
sys.path.insert(0, '.')
import numpy as np
import symboltz.jl_b200 as sb
ks = np.array([0.7, 30.0, 400.0])
for M in (sb.ΛCDM(lmax=10), sb.w0waCDM(lmax=10), sb.ΛCDM(lmax=10, nx=8), sb.ΛCDM(lmax=32)):
    prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
    ref = sb.solvept(prob, bg, ks, reltol=1e-10, abstol=1e-10)
    ref2 = sb.solvept(prob, bg, ks, reltol=1e-8, abstol=1e-8)
    scale = np.abs(ref.uend).max(axis=1, keepdims=True)
    print(M, "Rodas5P 1e-8 vs 1e-10:", (np.abs(ref2.uend - ref.uend) / scale).max(axis=1))
    for alg in ("KenCarp4", "TRBDF2"):
        for tol in (1e-5, 1e-7):
            s = sb.solvept(prob, bg, ks, reltol=tol, abstol=tol, alg=alg, warn=False)
            print("   ", alg, tol, "ok", s.success, "steps", s.stats[:, 0].tolist(), "err per mode", (np.abs(s.uend - ref.uend) / scale).max(axis=1))
