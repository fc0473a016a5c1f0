"""A/B of the fused source evaluation (sources formed inside the integrator, only S written) against the round-1 path (saved states
written to HBM, second kernel) at the bench size: per-stage CUDA-event times, identical bits of S and C_l, HBM bytes saved.
usage: python scripts/ab_fused.py [reps] [--prof]   (--prof: one fused plan.run() only, for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 5
prof = "--prof" in sys.argv
M = sb.ΛCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
ls = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))
jl = sb.SphericalBesselCache(ls, xcut=2e3 * bg.tau0 * 1.02)
res = {}
for fused in ((True,) if prof else (False, True)):
    plan = sb.CMBPlan(prob, bg, jl, direct=True, fused=fused)
    plan.upload(); plan.run(); plan.learn_schedule(); plan.run(); torch.cuda.synchronize()
    if prof:
        plan.run(); torch.cuda.synchronize(); break
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    for a, b, c in ev:
        a.record(); plan.solve(); b.record(); plan.los_cl(); c.record()
    torch.cuda.synchronize()
    if fused:  # how much a save point weighs in the schedule's cost model (units of one attempt)
        for sc in (0.0, 0.06, 0.12, 0.2, 0.3):
            plan.learn_schedule(plan.cost_model, save_cost=sc); plan.run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): plan.solve()
            e1.record(); torch.cuda.synchronize()
            print(f"   save_cost={sc}: solve {e0.elapsed_time(e1) / 3:.2f} ms", flush=True)
        plan.d_items = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        plan.solve(); e0.record()
        for _ in range(3): plan.solve()
        e1.record(); torch.cuda.synchronize()
        print(f"   queue: solve {e0.elapsed_time(e1) / 3:.2f} ms", flush=True)
    res[fused] = dict(solve_ms=np.median([a.elapsed_time(b) for a, b, c in ev]), los_cl_ms=np.median([b.elapsed_time(c) for a, b, c in ev]), S=plan.d_S.cpu().numpy(), Cl=plan.download(),
                      mem_MB=(0 if fused else plan.d_usave.numel() * 8 / 1e6))
    print(f"fused={fused}: solve (to S) {res[fused]['solve_ms']:.2f} ms, LOS+C_l {res[fused]['los_cl_ms']:.2f} ms, saved-state buffer {res[fused]['mem_MB']:.0f} MB", flush=True)
    del plan
    torch.cuda.empty_cache()
if not prof:
    a, b = res[True]["S"][:, :, :-1], res[False]["S"][:, :, :-1]
    print("S fused vs state path: max |dS| / max|S| per mode =", float((np.abs(a - b).max(axis=2) / np.abs(b).max(axis=2)).max()), " C_l max rel diff:", float(np.abs(res[True]["Cl"] / res[False]["Cl"] - 1).max()))
