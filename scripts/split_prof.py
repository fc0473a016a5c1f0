"""Cycle accounting of the split kernel's attempt by phase (build with -DSB_SPLIT_PROF into scripts/variants/prof.so first).  GPU box only."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
lib = C.CDLL(os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "scripts/variants/prof.so"))
prob.lib = lib
names = ["accept path tail (or reject + wait for the controller warp) -> attempt start", "barrier at the attempt start", "f0, dT after a move", "sb_factor (rows, eliminations on warp 0 | Z, K1, basis slots 1-5 on the others) + barrier",
         "3-column solve + barrier", "hub dots of Z, 2x2 inverse", "stages: hub dots + accumulation + barrier (x7)", "stages: f-evaluation + barrier (x7)", "stages: B-solve on warp 0 + barrier (x7)",
         "k8 correction + error norm", "pick-up of the controller warp's answer after an accepted step", "accept: dense output, state update, slot 5 -> slot 0",
         "  (in sb_factor) clear, row scatter, 2 barriers", "  (in sb_factor) phase-0 path eliminations", "  (in sb_factor) top: gather of the children's Schur terms", "  (in sb_factor) 5x5 pivoted Gauss-Jordan"]
ks = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
f = lambda k: min(1e-2 / k, 1e-4)
out = (C.c_ulonglong * 24)()
for label, kw in (("config 1 (100 modes, no save times)", dict(ptivini=f)), ):
    sb.solvept(prob, bg, ks, split=True, warn=False, **kw); lib.sbm_split_prof(out, 1)
    s = sb.solvept(prob, bg, ks, split=True, warn=False, **kw); lib.sbm_split_prof(out, 1)
    v = np.array(list(out), dtype=np.float64); att = v[20]; tot = v[:16].sum()
    print(f"{label}: {int(att)} attempts, {tot / att:.0f} cycles per attempt (thread 0 of the CTA)")
    for n, c in zip(names, v[:16]):
        print(f"  {c / att:8.0f} cycles  {100 * c / tot:5.1f} %  {n}")
