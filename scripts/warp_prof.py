"""Cycle accounting of the warp-per-mode kernel's attempt by phase for a lone warp per SM (build with -DSB_WARP_PROF into scripts/variants/wprof.so first).  GPU box only."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
lib = C.CDLL(os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "scripts/variants/wprof.so")); prob.lib = lib
names = ["accept path tail -> attempt start", "request of the table rows (slots 1-5)", "f0, dT after a move", "sb_factor", "Z, K1 set-up, 3-column solve, hub dots of Z, 2x2 inverse",
         "wait for the rows + Hermite sweep of slots 1-5", "stages: hub dots + accumulation (x7)", "stages: f-evaluation (x7)", "stages: B-solve (x7)", "k8 correction + error norm",
         "controller (log2 + exp2)", "accept: dense output, state update, slot 5 -> slot 0", "", "", "", "", "  (in the request, lane 1) fence, expect_tx, search of the stage time in the knot times", "  (in the request, lane 1) cp.async.bulk issue"]
f = lambda k: min(1e-2 / k, 1e-4)
out = (C.c_ulonglong * 24)()
for label, ks in (("config 1: 100 modes, one warp per SM (latency regime)", sb.loggrid(1e-4, 1.0, length=100) / sb.k0), ("6057 modes, 8 warps per SM (throughput regime)", np.linspace(0.01, 2000, 3 * 2019))):
    kw = dict(ptivini=f) if len(ks) == 100 else {}
    sb.solvept(prob, bg, ks, split=False, warn=False, **kw); lib.sbm_split_prof(out, 1)
    s = sb.solvept(prob, bg, ks, split=False, warn=False, **kw); lib.sbm_split_prof(out, 1)
    v = np.array(list(out), dtype=np.float64); att = v[20]; tot = v[:12].sum()
    print(f"{label}: {int(att)} attempts, {tot / att:.0f} cycles per attempt (lane 0 of every warp)")
    for n, c in [(n, c) for n, c in zip(names, v[:18]) if n]:
        print(f"  {c / att:8.0f} cycles  {100 * c / tot:5.1f} %  {n}")
