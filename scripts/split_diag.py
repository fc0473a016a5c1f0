"""Where do the split kernel (one CTA of SB_R warps per mode) and the warp-per-mode kernel part ways?  Compares the end states after
1, 2, 3, 10, 100 attempts (maxiters) and the full solve, bit for bit.  GPU box only."""
import sys
sys.path.insert(0, '.')
import numpy as np
import symboltz.jl_b200 as sb

for lmax, nx, w0wa in ((5, 4, False), (10, 4, False), (10, 4, True), (10, 8, False)):
    M = sb.w0waCDM(lmax=lmax, nx=nx) if w0wa else sb.ΛCDM(lmax=lmax, nx=nx); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
    ks = sb.loggrid(1e-4, 1.0, length=40) / sb.k0
    _, taus = sb.cmb_grids(bg)
    for mi in (1, 2, 3, 10, 100, 100000):
        a = sb.solvept(prob, bg, ks, split=False, maxiters=mi, warn=False)
        b = sb.solvept(prob, bg, ks, split=True, maxiters=mi, warn=False)
        d = np.abs(a.uend - b.uend) / (np.abs(a.uend) + 1e-300)
        bad = np.argwhere(~((a.uend == b.uend) | (np.isnan(a.uend) & np.isnan(b.uend))))
        print(f"{M} maxiters {mi}: stats equal {np.array_equal(a.stats, b.stats)}, differing entries {len(bad)} of {a.uend.size}, max rel {np.nanmax(d):.2e}, "
              f"first (mode, var) {bad[:6].tolist()}", flush=True)
    for nS in (2, 3):
        kw = dict(saveat=taus[::6], sources=dict(nS=nS, scale_k=True))
        a = sb.solvept(prob, bg, ks, split=False, warn=False, **kw)
        b = sb.solvept(prob, bg, ks, split=True, warn=False, **kw)
        A_, B_ = a.d_S.cpu().numpy(), b.d_S.cpu().numpy()
        print(f"{M} nS {nS}: usave equal {np.array_equal(a.usave, b.usave, equal_nan=True)}, S equal {np.array_equal(A_, B_, equal_nan=True)}, "
              f"max |dS|/max|S| {np.nanmax(np.abs(A_ - B_)) / np.nanmax(np.abs(A_)):.2e}", flush=True)
