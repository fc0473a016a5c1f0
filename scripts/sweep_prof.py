"""Where the time of a config-4 sweep goes: host backgrounds, per-chunk host preparation, integrator launch, attempts per second."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, warnings
import concurrent.futures as cf
import symboltz.jl_b200 as sb
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
M = sb.w0waCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
names = ["h", "Omega_c", "Omega_b", "ln_As1e10", "ns", "w0", "wa"]
lo = np.array([0.6, 0.10, 0.020, 2.9, 0.92, -1.2, -0.3]); hi = np.array([0.8, 0.14, 0.025, 3.2, 1.0, -0.8, 0.3])
rng = np.random.default_rng(0)
u = (rng.permuted(np.tile(np.arange(nc), (7, 1)), axis=1).T + rng.random((nc, 7))) / nc
th = lo + (hi - lo) * u
th[:, 1] /= th[:, 0] ** 2; th[:, 2] /= th[:, 0] ** 2
ks = sb.loggrid(1e-4, 1.0, length=256) / sb.k0
upd = sb.parameter_updater(prob, names)
def host(t):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore"); p = upd(t); return sb.solvebg(p)
t0 = time.time()
with cf.ThreadPoolExecutor(os.cpu_count()) as pool: bgs = list(pool.map(host, th))
print(f"host backgrounds: {time.time()-t0:.3f} s for {nc} on {os.cpu_count()} threads; knots {np.mean([len(b.t) for b in bgs]):.0f}")
f = lambda k: min(1e-2 / k, 1e-4)
sb.solvept_batch(bgs[:2], ks, ptivini=f); torch.cuda.synchronize()
for b in bgs[:2]: b._dev = None
for c0 in range(0, nc, chunk):
    g = bgs[c0:c0 + chunk]
    t1 = time.time()
    for b in g: b.device()
    torch.cuda.synchronize(); t2 = time.time()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); batch = sb.solvept_batch(g, ks, ptivini=f); e.record(); t3 = time.time(); torch.cuda.synchronize()
    st = batch.d_stats.cpu().numpy(); att = st[:, 0] + st[:, 1]
    rcs = batch.d_retcode.cpu().numpy()
    print(f"chunk {c0//chunk}: device() {1e3*(t2-t1):.1f} ms | solvept_batch host {1e3*(t3-t2):.1f} ms, GPU {a.elapsed_time(e):.1f} ms | attempts {att.sum()/1e6:.2f} M max {att.max()} -> {att.sum()/a.elapsed_time(e)/1e3:.1f} M/s | fails {np.unique(rcs[rcs!=0], return_counts=True)} worst k {np.concatenate([ks]*len(g))[rcs!=0][:4]}")
    for b in g: b._dev = None
