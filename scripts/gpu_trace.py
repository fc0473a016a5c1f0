import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=5)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
k = float(sys.argv[1]) if len(sys.argv) > 1 else 1000.0
sol = sb.solvept(prob, bg, [k], trace=4000)
np.save("gpurun_out/trace_gpu.npy", sol.trace); np.savez("gpurun_out/bg_knots.npz", t=bg.t, y=bg.y, dy=bg.dy, tau0=bg.tau0, kappa0=bg.kappa0)
code = """
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import sbref
d = np.load('gpurun_out/bg_knots.npz')
obg = sbref.Background.from_knots(sbref.planck18(lmax=5), d['t'], d['y'], d['dy'], float(d['tau0']), float(d['kappa0']))
sbref.solvept(obg, [%r], nthreads=1)
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), k)
r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SBO_TRACE="1"), capture_output=True, text=True)
tr = np.array([[float(x) for x in l.split()[2:]] for l in r.stderr.splitlines() if l.startswith("TRACE")])
g = sol.trace[: len(tr)]
n = min(len(tr), (g[:, 1] != 0).sum())
rel = np.abs(g[:n] / tr[:n] - 1)
print("attempts oracle", len(tr), "gpu", (sol.trace[:, 1] != 0).sum())
for i in list(range(0, 12)) + list(range(12, n, max(1, n // 25))):
    print(i, "t %.6e dt %.6e EEst %.6e | rel diff t %.1e dt %.1e EEst %.1e" % (tr[i, 0], tr[i, 1], tr[i, 2], rel[i, 0], rel[i, 1], rel[i, 2]))
bad = np.nonzero(rel[:, 1] > 1e-3)[0]
print("first attempt with dt rel diff > 1e-3:", bad[:1], "EEst there", tr[bad[:1], 2] if len(bad) else None, g[bad[:1], 2] if len(bad) else None)
