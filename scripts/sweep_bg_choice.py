"""Config-4 sweep (512 w0waCDM cosmologies x 256 k on one GPU): host background pool with 16 and with 4 threads (what a rank has at N = 8 on a 32-core box)
against the batched device background solver.  GPU box only."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import config4_thetas
Mw = sb.w0waCDM(lmax=10); probw = sb.CosmologyProblem(Mw, sb.parameters_Planck18(Mw))
names, th = config4_thetas(512)
ks = sb.loggrid(1e-4, 1.0, length=256) / sb.k0
ref = None
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    sb.spectrum_matter_sweep(probw, names, th[:4], ks)
    sb.spectrum_matter_sweep(probw, names, th[:4], ks, background="device")
    for label, kw in (("host pool, 16 threads", dict(nthreads=16)), ("host pool, 4 threads", dict(nthreads=4)), ("device, one thread per cosmology", dict(background="device")),
                      ("host pool, 4 threads", dict(nthreads=4)), ("device, one thread per cosmology", dict(background="device"))):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        P, info = sb.spectrum_matter_sweep(probw, names, th, ks, chunk=32, return_info=True, **kw)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        if ref is None: ref = P
        ok = np.isfinite(P).all(axis=1) & np.isfinite(ref).all(axis=1)
        print(f"{label:36s}: {t:.2f} s = {512 * 256 / t / 1e3:.1f} k modes/s; background failures {info['background_failures']}, mode failures {info['mode_failures']}, finite rows {int(np.isfinite(P).all(axis=1).sum())}; "
              f"max |P/P_host16 - 1| over common finite rows {np.abs(P[ok] / ref[ok] - 1).max():.1e}", flush=True)
