"""Scratch GPU check: sources, LOS, C_l vs oracle and vs CLASS golden (lmax=5)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
from oracle import sbref

def timed(f):
    torch.cuda.synchronize(); t = time.time(); r = f(); torch.cuda.synchronize(); return r, time.time() - t

M = sb.ΛCDM(lmax=5)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
obg = sbref.Background.from_knots(sbref.planck18(lmax=5), bg.t, bg.y, bg.dy, bg.tau0, bg.kappa0)
ls = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))
# Bessel table
jl, tj = timed(lambda: sb.SphericalBesselCache(ls, xcut=2e3 * bg.tau0 * 1.001))
print("bessel table", jl.y.shape, "%.3fs" % tj)
t = time.time(); ojl = sbref.SphericalBesselCache(ls, xcut=2e3 * bg.tau0 * 1.001); print("oracle bessel %.1fs" % (time.time() - t), ojl.y.shape)
y = jl.y.cpu().numpy(); dy = jl.dy.cpu().numpy()
n = min(y.shape[0], ojl.y.shape[1])
print("bessel max abs diff y", np.abs(y[:n].T - ojl.y[:, :n]).max(), "dy", np.abs(dy[:n].T - ojl.dy[:, :n]).max())
# sources on coarse nodes
kint = sb.ChebyshevInterpolator(1e-2, 2e3, 60)
ks_fine, taus = sb.cmb_grids(bg)
S, ts = timed(lambda: sb.source_grid(prob, taus, kint.xs, bg))
print("source_grid 61 modes %.3fs" % ts, "ok", S.sol.success, "steps", S.sol.stats[:, 0].sum())
osol = sbref.solvept(obg, kint.xs, saveat=taus)
oS = sbref.sources(obg, kint.xs, taus, osol["usave"])
Sg = S.dS.cpu().numpy()  # [nk][2][nt]
oST = kint.xs[:, None] * oS[:, :, 0]; oSE = kint.xs[:, None] ** 2 * oS[:, :, 1]
sc = np.abs(oST).max(axis=1, keepdims=True)
print("usave rel diff", np.nanmax(np.abs(S.sol.usave - osol["usave"])) / np.nanmax(np.abs(osol["usave"])))
print("ST rel-to-max diff", (np.abs(Sg[:, 0, :] - oST) / sc).max(), "SE", (np.abs(Sg[:, 1, :-1] - oSE[:, :-1]) / np.abs(oSE[:, :-1]).max(axis=1, keepdims=True)).max())
for direct in (False, True):
    for rep in range(2):
        (Cl, ex), tc = timed(lambda: sb.spectrum_cmb(["TT", "EE", "TE"], prob, jl, normalization="Dl", bgsol=bg, direct=direct, return_all=True))
    print("spectrum_cmb direct=%s: %.3fs" % (direct, tc))
    t = time.time(); oCl, oex = sbref.spectrum_cmb(["TT", "EE", "TE"], obg, ojl, normalization="Dl", direct=direct, return_all=True); to = time.time() - t
    th = ex["theta"].cpu().numpy()
    print("  oracle %.1fs; Theta_T rel-to-max diff" % to, np.abs(th[0].T - oex["ThT"]).max() / np.abs(oex["ThT"]).max(), "Theta_E", np.abs(th[1].T - oex["ThE"]).max() / np.abs(oex["ThE"]).max())
    rel = np.abs(Cl / oCl - 1)
    print("  C_l max rel diff TT/EE/TE:", rel.max(axis=0), "(TE relative to max:", np.abs(Cl[:, 2] - oCl[:, 2]).max() / np.abs(oCl[:, 2]).max(), ")")
    if not direct:
        d = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "class_Cl.dat"))
        Dl_all = sb.spline_ls(Cl, ls, d[:, 0])
        print("  vs CLASS golden normwise TT %.3e EE %.3e" % (np.linalg.norm(Dl_all[:, 0] - d[:, 1]) / np.linalg.norm(d[:, 1]), np.linalg.norm(Dl_all[:, 1] - d[:, 2]) / np.linalg.norm(d[:, 2])))
