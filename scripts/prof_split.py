"""For ncu: config 1 (100 modes) once with the warp-per-mode kernel and once with the split kernel (one CTA of SB_R row warps + a controller warp per mode),
each after a warm-up launch:  ncu --set full -k regex:sb_integrate -s 2 -c 2 ... python scripts/prof_split.py   (launches 1, 2 are the warm-ups)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
f = lambda k: min(1e-2 / k, 1e-4)
for split in (False, True):
    sb.solvept(prob, bg, ks, ptivini=f, split=split, warn=False)
torch.cuda.synchronize()
for split in (False, True):
    s = sb.solvept(prob, bg, ks, ptivini=f, split=split, warn=False)
    torch.cuda.synchronize()
    print("split" if split else "warp per mode", int((s.stats[:, 0] + s.stats[:, 1]).sum()), "attempts in the launch")
