"""Scratch GPU check (run under gpurun): P(k) parity vs oracle + golden, timing."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
from oracle import sbref

def timed(f):
    torch.cuda.synchronize(); t = time.time(); r = f(); torch.cuda.synchronize(); return r, time.time() - t

out = {}
M = sb.ΛCDM(lmax=5)
pars = sb.parameters_Planck18(M)
prob = sb.CosmologyProblem(M, pars)
bg = sb.solvebg(prob)
d = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "class_Pk.dat"))
ks, Pc = d[:, 0], d[:, 1]
(P, sol), t1 = timed(lambda: sb.spectrum_matter(prob, ks, bgsol=bg, return_solution=True))
(P, sol), t2 = timed(lambda: sb.spectrum_matter(prob, ks, bgsol=bg, return_solution=True))
print("lmax5 516 modes: first call %.3fs second %.3fs" % (t1, t2), "grid", sol.grid, "success", sol.success)
st = sol.stats
print("steps acc/rej total", st[:, 0].sum(), st[:, 1].sum(), "max", st[:, 0].max())
print("vs golden normwise", np.linalg.norm(P - Pc) / np.linalg.norm(Pc), "max rel (k>5th)", np.abs(P / Pc - 1)[5:].max())
obg = sbref.Background.from_knots(sbref.planck18(lmax=5), bg.t, bg.y, bg.dy, bg.tau0, bg.kappa0)
t = time.time(); Po, osol = sbref.spectrum_matter(obg, ks); to = time.time() - t
rel = np.abs(P / Po - 1)
print("oracle time %.2fs; GPU vs oracle max rel %.3e median %.3e at k=%g" % (to, rel.max(), np.median(rel), ks[rel.argmax()]))
print("steps gpu vs oracle (first 10 / last 10):", st[:10, 0], osol["stats"][:10, 0], st[-10:, 0], osol["stats"][-10:, 0])
print("uend max rel diff", np.abs(sol.uend - osol["uend"]).max() / np.abs(osol["uend"]).max())
out["pk_lmax5"] = dict(t=t2, oracle_t=to, maxrel=float(rel.max()), normwise_golden=float(np.linalg.norm(P - Pc) / np.linalg.norm(Pc)))

M = sb.ΛCDM(lmax=10)
prob10 = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg10 = sb.solvebg(prob10)
ks100 = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
for rep in range(3):
    (P10, sol10), t3 = timed(lambda: sb.spectrum_matter(prob10, ks100, bgsol=bg10, return_solution=True))
    print("lmax10 100 modes: %.4fs" % t3, "steps max", sol10.stats[:, 0].max(), "total", sol10.stats[:, 0].sum())
ks2000 = np.linspace(0.01, 2000, 2020)
for rep in range(2):
    s2, t4 = timed(lambda: sb.solvept(prob10, bg10, ks2000))
    print("lmax10 2020 modes to tau0 from 1e-6: %.4fs" % t4, "grid", s2.grid, "ok", s2.success, "steps total", s2.stats[:, 0].sum(), "max", s2.stats[:, 0].max(), "modes/s", 2020 / t4)
out["t_100"] = t3; out["t_2020"] = t4
json.dump(out, open("gpurun_out/gpu_check.json", "w"))
