"""Same-box A/B of engine builds on the config-4 sweep (w0waCDM hypercube x 256 modes): python scripts/sweep_ab.py NC a.so b.so ...
Each variant must be a build of the l10_x4_w0wa engine; it is slipped into the library cache under the official path."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
from symboltz.jl_b200 import api, build
nc = int(sys.argv[1])
so, _ = build.build_model(10, 4, True)
names = ["h", "Omega_c", "Omega_b", "ln_As1e10", "ns", "w0", "wa"]
lo = np.array([0.6, 0.10, 0.020, 2.9, 0.92, -1.2, -0.3]); hi = np.array([0.8, 0.14, 0.025, 3.2, 1.0, -0.8, 0.3])
rng = np.random.default_rng(0)
u = (rng.permuted(np.tile(np.arange(nc), (7, 1)), axis=1).T + rng.random((nc, 7))) / nc
th = lo + (hi - lo) * u
th[:, 1] /= th[:, 0] ** 2; th[:, 2] /= th[:, 0] ** 2
ks = sb.loggrid(1e-4, 1.0, length=256) / sb.k0
ref = None
for rnd in range(2):
    for v in sys.argv[2:]:
        api._libs[so] = C.CDLL(os.path.abspath(v))
        M = sb.w0waCDM(lmax=10)
        prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
        prob.lib.sbm_resident_warps_batch.restype = C.c_int
        sb.spectrum_matter_sweep(prob, names, th[:8], ks)
        torch.cuda.synchronize(); t = time.time(); P, info = sb.spectrum_matter_sweep(prob, names, th, ks, chunk=64, return_info=True); dt = time.time() - t
        if ref is None: ref = P
        ok = np.isfinite(P) & np.isfinite(ref)
        print(f"round {rnd} {os.path.basename(v):22s} batch warps {prob.lib.sbm_resident_warps_batch():5d}: {dt:.3f} s -> {nc*256/dt:.0f} k-modes/s {info} max rel diff vs first {np.abs(P[ok]/ref[ok]-1).max():.1e}")
