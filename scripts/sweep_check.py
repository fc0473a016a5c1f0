"""Config-4 style sweep (w0waCDM Latin hypercube, 256 log-spaced k): correctness vs individual calls and throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = sb.w0waCDM(lmax=10)
pars = sb.parameters_Planck18(M)
prob = sb.CosmologyProblem(M, pars)
names = ["h", "Omega_c", "Omega_b", "ln_As1e10", "ns", "w0", "wa"]
lo = np.array([0.6, 0.10, 0.020, 2.9, 0.92, -1.2, -0.3]); hi = np.array([0.8, 0.14, 0.025, 3.2, 1.0, -0.8, 0.3])
rng = np.random.default_rng(0)
u = (rng.permuted(np.tile(np.arange(nc), (7, 1)), axis=1).T + rng.random((nc, 7))) / nc  # Latin hypercube
th = lo + (hi - lo) * u
th[:, 1] /= th[:, 0] ** 2; th[:, 2] /= th[:, 0] ** 2   # Ωc h², Ωb h² -> Ωc, Ωb
ks = sb.loggrid(1e-4, 1.0, length=256) / sb.k0
sb.spectrum_matter_sweep(prob, names, th[:8], ks)  # warm-up (library load, allocator)
for chunk in ([int(a) for a in sys.argv[2:]] or [32]):
    torch.cuda.synchronize(); t = time.time(); P, info = sb.spectrum_matter_sweep(prob, names, th, ks, chunk=chunk, return_info=True); dt = time.time() - t
    if os.environ.get("SB_TIMELINE"):
        _, inf = sb.spectrum_matter_sweep(prob, names, th, ks, chunk=chunk, return_info="timeline", nslots=int(os.environ.get("SB_SLOTS", 2)))
        for ev in inf["timeline"]: print("   %-7s slot %d  %7.1f -> %7.1f ms" % (ev[0], ev[1], 1e3 * ev[2], 1e3 * ev[3]))
    print(f"{nc} cosmologies x {len(ks)} modes, {chunk} cosmologies per launch: {dt:.2f} s wall (host backgrounds included, {os.cpu_count()} cores) -> {nc*len(ks)/dt:.0f} k-modes/s; {info}")
upd = sb.parameter_updater(prob, names)
for i in (0, nc // 2):
    Pi = sb.spectrum_matter(upd(th[i]), ks)
    print("cosmology", i, "sweep vs single call max rel diff", np.abs(P[i] / Pi - 1).max())
t = time.time()
for i in range(4): sb.solvebg(upd(th[i]))
print("host background solve: %.3f s each" % ((time.time() - t) / 4))
