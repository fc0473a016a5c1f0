"""Which (cosmology, k) pairs of the config-4 Latin hypercube fail, and with what return code."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, warnings, json
import concurrent.futures as cf
import symboltz.jl_b200 as sb
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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
with cf.ThreadPoolExecutor(os.cpu_count()) as pool: bgs = list(pool.map(host, th))
f = lambda k: min(1e-2 / k, 1e-4)
out = []
for c0 in range(0, nc, 64):
    g = bgs[c0:c0 + 64]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore"); batch = sb.solvept_batch(g, ks, ptivini=f)
    rc = batch.d_retcode.cpu().numpy().reshape(len(g), -1); st = batch.d_stats.cpu().numpy().reshape(len(g), len(ks), 4)
    for i, j in zip(*np.nonzero(rc)):
        out.append(dict(cosmo=int(c0 + i), theta=th[c0 + i].tolist(), k=float(ks[j]), ik=int(j), rc=int(rc[i, j]), stats=st[i, j].tolist(), tau0=g[i].tau0))
    att = st[..., 0] + st[..., 1]
    print(f"chunk {c0//64}: max attempts per cosmology: min {att.max(1).min()} median {np.median(att.max(1))} max {att.max(1).max()}; argmax cosmology theta {th[c0 + att.max(1).argmax()]}")
    for b in g: b._dev = None
print(len(out), "failures")
for o in out[:40]: print(o)
json.dump(out, open("gpurun_out/sweep_failures.json", "w"))
