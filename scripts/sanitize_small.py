"""Tiny workload for compute-sanitizer (memcheck / racecheck): a few modes through the integrator (queue, static schedule and batched launch),
the source kernels and the LOS/C_l kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
FAST = os.environ.get("SB_SANITIZE_FAST") == "1"  # memcheck in ≈2 minutes: one model, a coarse k-interpolator for the C_l call
for M in ((sb.ΛCDM(lmax=5),) if FAST else (sb.ΛCDM(lmax=5), sb.ΛCDM(lmax=10))):
    prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
    bg = sb.solvebg(prob)
    ks = np.array([0.5, 20.0, 300.0, 1500.0])
    P = sb.spectrum_matter(prob, ks, bgsol=bg)
    cost = sb.ModeCostModel(ks, np.array([300.0, 500.0, 1200.0, 2500.0]), nknots=3) if hasattr(sb, "ModeCostModel") else None
    f = lambda k: min(1e-2 / k, 1e-4)
    s1 = sb.solvept(prob, bg, ks, ptivini=f)
    s2 = sb.solvept(prob, bg, ks, ptivini=f, cost=np.array([300.0, 500.0, 1200.0, 2500.0]))
    b = sb.solvept_batch([bg, bg], ks, ptivini=f)
    torch.cuda.synchronize()
    print(M, "P(k)", P, "sched == queue", np.array_equal(s1.uend, s2.uend), "batch == queue", np.array_equal(b.sols[1].uend, s1.uend))
ls = np.array([10, 100, 1000])
jl = sb.SphericalBesselCache(ls, xmax=2.1e3 * bg.tau0)
print(sb.spectrum_cmb(["TT", "EE"], prob, jl, bgsol=bg, kinterp=sb.ChebyshevInterpolator(1e-2, 400.0, 10) if FAST else None)[:, 0])
d = sb.solvebg_batch([prob, prob])
print("device bg tau0", d[0].tau0, bg.tau0)
# round 2: fused sources (queue, static schedule with parking, batched) and the lockstep parameter lanes
taus = np.geomspace(1e-3, bg.tau0, 21); taus[-1] = bg.t[-1]
a = sb.source_grid(prob, taus, ks, bg)
b = sb.source_grid(prob, taus, ks, bg, cost=np.array([300.0, 500.0, 1200.0, 2500.0]))
c = sb.solvept_batch([bg, bg], ks, saveat=[taus, taus], sources=dict(nS=3, scale_k=True), keep_states=False)
torch.cuda.synchronize()
print("fused: sched == queue", np.array_equal(a.dS.cpu().numpy(), b.dS.cpu().numpy(), equal_nan=True), "finite", bool(np.isfinite(c.sols[1].d_S.cpu().numpy()[:, :, :-1]).all()))
J = sb.sensitivity_matter(prob, ["Omega_c", "Omega_b", "ns"], ks[:3])
ls2 = np.array([30, 300])
Jc = sb.sensitivity_cmb("TT", prob, ["Omega_c", "h"], sb.SphericalBesselCache(ls2, xmax=2.1e3 * bg.tau0), kinterp=sb.ChebyshevInterpolator(1e-2, 400.0, 12))
torch.cuda.synchronize()
print("lanes dlnP/dlnθ", J[0], "dlnD_l/dlnθ", Jc[0])
