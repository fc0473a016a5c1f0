"""Cost of a save point (dense output + source evaluation inside the integrator) as a function of how many there are: 2019 modes from the queue, with 0, 75, 300, 1200 save
times.  If the per-point cost falls steeply with the density of save times, the cost is the excursion into cold code (instruction cache), not the arithmetic.  GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks, taus = sb.cmb_grids(bg)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
t0 = timed(lambda: sb.solvept(prob, bg, ks, warn=False))
print(f"no save times: {t0:.2f} ms")
for n in (75, 300, 1200, 4800):
    tt = np.interp(np.linspace(0, 1, n), np.linspace(0, 1, len(taus)), taus)
    t = timed(lambda: sb.solvept(prob, bg, ks, saveat=tt, sources=dict(nS=2, scale_k=True), keep_states=False, warn=False))
    per = (t - t0) * 1e3 * 1184 / (len(ks) * n)  # us of warp time per save point (1184 resident warps)
    print(f"{n:5d} save times: {t:.2f} ms (+{t - t0:.2f} ms) = {per:.2f} us of a warp's time per save point", flush=True)
