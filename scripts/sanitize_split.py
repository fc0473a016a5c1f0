"""Tiny workload for compute-sanitizer on the split kernel (one CTA of SB_R row warps + a controller warp per mode): a few cheap modes, bounded attempts,
with and without save times / fused sources, for SB_R = 2 (lmax = 5), 3 (lmax = 10) and 4 (nx = 8).  Compares with the warp-per-mode kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
models = (sb.ΛCDM(lmax=5),) if os.environ.get("SB_SANITIZE_FAST") == "1" else (sb.ΛCDM(lmax=5), sb.ΛCDM(lmax=10), sb.ΛCDM(lmax=10, nx=8))
for M in models:
    prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
    ks = np.array([0.0, 0.5, 5.0, 40.0])
    taus = np.geomspace(1e-3, bg.tau0, 9); taus[-1] = bg.t[-1]
    for kw in (dict(maxiters=120), dict(maxiters=120, saveat=taus, sources=dict(nS=3, scale_k=True))):
        a = sb.solvept(prob, bg, ks, split=False, warn=False, **kw)
        b = sb.solvept(prob, bg, ks, split=True, warn=False, **kw)
        torch.cuda.synchronize()
        same = np.array_equal(a.uend, b.uend, equal_nan=True) and np.array_equal(a.stats, b.stats) and np.array_equal(a.retcode, b.retcode)
        if "saveat" in kw:
            same = same and np.array_equal(a.usave, b.usave, equal_nan=True) and np.array_equal(a.d_S.cpu().numpy(), b.d_S.cpu().numpy(), equal_nan=True)
        print(M, sorted(kw), "split == warp:", same, "retcodes", b.retcode.tolist(), flush=True)
