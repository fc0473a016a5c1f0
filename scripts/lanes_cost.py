"""Cost of the lockstep parameter lanes (BASELINE config 5) against one primal solve, CUDA-event timed:
default C_l path (61 Chebyshev nodes) and the direct path (all ~2020 fine modes), lanes = primal + (h, Omega_c, Omega_b)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg0 = sb.solvebg(prob)
names, delta = ["h", "Omega_c", "Omega_b"], 1e-5
_, probs, _steps = sb.api._lane_problems(prob, names, delta)
bgs = sb.api._lane_backgrounds(probs, bg0)
ks_fine, taus = sb.cmb_grids(bg0)
cg = sb.cosgrid(0.0, 1.0, length=300)
i0 = int(np.searchsorted(bg0.t, 1e-2, side="left"))
saves = []
for b in bgs:
    tj = b.t[i0] + (b.t[-1] - b.t[i0]) * cg; tj[-1] = b.t[-1]; saves.append(tj)
def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
src = dict(nS=2, scale_k=True)
for label, ks in (("61 Chebyshev nodes", sb.ChebyshevInterpolator(1e-2, 2e3, 60).xs), ("2019 fine modes", ks_fine)):
    t1 = timed(lambda: sb.solvept(prob, bgs[0], ks, saveat=saves[0], sources=src, keep_states=False, warn=False))
    tb = timed(lambda: sb.solvept_batch(bgs, ks, saveat=saves, sources=src, keep_states=False))
    tl = timed(lambda: sb.solvept_lanes(bgs, ks, [0.0] + [1 / delta] * 3, saveat=saves, sources=src))
    t0 = timed(lambda: sb.solvept_lanes(bgs, ks, [0.0] * 4, saveat=saves, sources=src))  # partials left out of the error norm: the primal's steps
    sols = sb.solvept_lanes(bgs, ks, [0.0] + [1 / delta] * 3, saveat=saves, sources=src)
    one = sb.solvept(prob, bgs[0], ks, saveat=saves[0], sources=src, keep_states=False, warn=False)
    a1, al = (one.stats[:, 0] + one.stats[:, 1]).sum(), (sols[0].stats[:, 0] + sols[0].stats[:, 1]).sum()
    print(f"{label}: primal solve {t1:.1f} ms ({a1} attempts) | 4 independent cosmologies in one batched launch {tb:.1f} ms | 4 lockstep lanes {tl:.1f} ms ({al} attempts per lane) -> {tl / t1:.2f}x the primal (1 + p = 4); with the primal's own steps (partials not in the norm) {t0:.1f} ms = {t0 / t1:.2f}x", flush=True)
