"""Same-box A/B of the integrator launch at the bench size: atomic queue (descending k) vs the static preemptive schedule
(cost model learnt on this cosmology, and transferred from a different cosmology)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10)
ls = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))
def mk(mod):
    pars = sb.parameters_Planck18(M)
    for k_, f in mod.items(): pars[k_] *= f
    prob = sb.CosmologyProblem(M, pars); bg = sb.solvebg(prob)
    jl = sb.SphericalBesselCache(ls, xcut=2e3 * bg.tau0 * 1.001)
    return sb.CMBPlan(prob, bg, jl)
def tm(plan, n=5):
    plan.solve(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        a.record(); plan.solve(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts), plan.d_uend.cpu().numpy().copy(), plan.d_usave.cpu().numpy().copy(), plan.d_stats.cpu().numpy().copy()
A = mk({}); A.upload()
tq, u0, us0, st0 = tm(A)
model = A.learn_schedule()
ts, u1, us1, st1 = tm(A)
print(f"planck: queue {tq:.2f} ms | static (own model) {ts:.2f} ms  ({100*(ts/tq-1):+.1f} %)  identical: uend {np.array_equal(u0,u1)} usave {np.array_equal(us0,us1,equal_nan=True)} stats {np.array_equal(st0,st1)} ok {(A.d_ret.cpu().numpy()==0).all()}")
att = st0[:, 0] + st0[:, 1]
print(f"  attempts sum {att.sum()} -> ideal {att.sum()/A.nlists:.0f} per warp over {A.nlists} lists; T = {A.sched_T:.0f}; items {len(A.d_items)}")
B = mk({"Omega_c": 1.10, "h": 0.95, "Omega_b": 1.05}); B.upload()
tq2, v0, vs0, vt0 = tm(B)
B.learn_schedule(model)  # cost model from the OTHER cosmology
ts2, v1, vs1, vt1 = tm(B)
print(f"variant cosmology ({B.nk} modes): queue {tq2:.2f} ms | static (planck's model) {ts2:.2f} ms ({100*(ts2/tq2-1):+.1f} %) identical: {np.array_equal(v0,v1)} {np.array_equal(vs0,vs1,equal_nan=True)} {np.array_equal(vt0,vt1)}")
B.learn_schedule()
ts3, *_ = tm(B)
print(f"  static (own model) {ts3:.2f} ms")
