"""Batched device background solve (sbm_solvebg_batch) vs the host solver: agreement, kernel time, and the config-4 sweep with
host vs device backgrounds at several host thread counts."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M = sb.w0waCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
names = ["h", "Omega_c", "Omega_b", "ln_As1e10", "ns", "w0", "wa"]
lo = np.array([0.6, 0.10, 0.020, 2.9, 0.92, -1.2, -0.3]); hi = np.array([0.8, 0.14, 0.025, 3.2, 1.0, -0.8, 0.3])
rng = np.random.default_rng(0)
u = (rng.permuted(np.tile(np.arange(nc), (7, 1)), axis=1).T + rng.random((nc, 7))) / nc
th = lo + (hi - lo) * u
th[:, 1] /= th[:, 0] ** 2; th[:, 2] /= th[:, 0] ** 2
upd = sb.parameter_updater(prob, names)
t = time.time(); probs = [upd(x) for x in th]; print(f"{nc} CosmologyProblem updates: {time.time()-t:.3f} s")
sb.solvebg_batch(probs[:4]); torch.cuda.synchronize()
for n in sorted({32, 128, nc}):
    t = time.time(); bgs = sb.solvebg_batch(probs[:n]); dtw = time.time() - t
    # kernel alone
    dP = torch.from_numpy(np.stack([p.P for p in probs[:n]])).cuda(); cap = 4096
    bufs = [torch.empty((n, cap), dtype=torch.float64, device="cuda"), torch.empty((n, cap, 5), dtype=torch.float64, device="cuda"), torch.empty((n, cap, 5), dtype=torch.float64, device="cuda"),
            torch.empty((n, 8), dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")]
    import ctypes as C
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    prob.lib.sbm_solvebg_batch(C.c_int(n), C.c_void_p(dP.data_ptr()), C.c_double(1e-6), C.c_double(100.0), C.c_double(1e-7), C.c_double(1e-7), C.c_int(cap), *[C.c_void_p(x.data_ptr()) for x in bufs], C.c_void_p(torch.cuda.current_stream().cuda_stream))
    b.record(); torch.cuda.synchronize()
    nbs = np.array([len(x.t) for x in bgs])
    print(f"solvebg_batch n={n}: kernel {a.elapsed_time(b):.1f} ms, call incl. download {1e3*dtw:.1f} ms; knots min/mean/max {nbs.min()}/{nbs.mean():.0f}/{nbs.max()}; steps/cosmology {np.mean([x.naccept + x.nreject for x in bgs]):.0f}; all ok {all(x.success for x in bgs)}")
t = time.time()
with warnings.catch_warnings():
    warnings.simplefilter("ignore"); hs = [sb.solvebg(p) for p in probs[:16]]
th_host = (time.time() - t) / 16
print(f"host solvebg: {1e3*th_host:.2f} ms per cosmology per core ({os.cpu_count()} cores on this box)")
dt0 = max(abs(d.tau0 / h.tau0 - 1) for d, h in zip(bgs, hs)); dk = max(abs(d.kappa0 / h.kappa0 - 1) for d, h in zip(bgs, hs))
print(f"device vs host: tau0 max rel {dt0:.2e}, kappa0 max rel {dk:.2e}, knots {[len(d.t) - len(h.t) for d, h in zip(bgs[:8], hs[:8])]}")
ks = sb.loggrid(1e-4, 1.0, length=256) / sb.k0
sb.spectrum_matter_sweep(prob, names, th[:8], ks)
for bgmode, nthr in (("host", None), ("host", 2), ("host", 1), ("device", 1)):
    torch.cuda.synchronize(); t = time.time()
    P, info = sb.spectrum_matter_sweep(prob, names, th, ks, chunk=64, return_info=True, background=bgmode, nthreads=nthr); dt = time.time() - t
    print(f"sweep {nc} x 256 modes, backgrounds on {bgmode:6s} ({nthr or os.cpu_count()} host threads): {dt:.2f} s -> {nc*256/dt:.0f} k-modes/s {info}")
    if bgmode == "host" and nthr is None: P0 = P
print("P(k) device-background vs host-background sweep: max rel diff", np.nanmax(np.abs(P / P0 - 1)))
