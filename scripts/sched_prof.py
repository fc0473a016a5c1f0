"""Where does a statically scheduled launch lose time?  Per-list wall clock (end time, wait for parked modes) and per-mode time vs attempts,
from a -DSB_TIME_PROF build (scripts/variants/tprof.so).  GPU box only."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks, taus = sb.cmb_grids(bg)
lib = C.CDLL(os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "scripts/variants/tprof.so")); prob.lib = lib
L = (C.c_ulonglong * (4096 * 3))(); Mo = (C.c_ulonglong * 8192)()
def prof():
    lib.sbm_time_prof(L, Mo, 1)
    return np.array(L, dtype=np.float64).reshape(-1, 3), np.array(Mo, dtype=np.float64)
kw = dict(saveat=taus, sources=dict(nS=2, scale_k=True), keep_states=False, warn=False)
for label, kws in (("no save times", dict(warn=False)), ("300 save times + sources", kw)):
    s1 = sb.solvept(prob, bg, ks, **kws); prof()
    s1 = sb.solvept(prob, bg, ks, **kws); l, m = prof()
    att, acc, rej = (s1.stats[:, 0] + s1.stats[:, 1]).astype(np.float64), s1.stats[:, 0].astype(np.float64), s1.stats[:, 1].astype(np.float64)
    tm = m[:len(ks)] * 1e-3  # us per mode under the queue (all warps busy most of the time)
    A = np.stack([acc, rej, np.ones_like(acc)], axis=1)
    coef, *_ = np.linalg.lstsq(A, tm, rcond=None)
    res = tm - A @ coef
    print(f"[{label}] queue: per-mode time = {coef[0]:.2f} us x accepted + {coef[1]:.2f} us x rejected + {coef[2]:.0f} us; rms residual {np.sqrt((res**2).mean()):.0f} us of mean {tm.mean():.0f} us; "
          f"time/attempt: min {np.min(tm / att):.2f} median {np.median(tm / att):.2f} max {np.max(tm / att):.2f} us")
    nl = int((l[:, 1] > 0).sum()); t0 = l[:nl, 0].min()
    print(f"   queue: {nl} warps, end times (ms after the first start): min {(l[:nl, 1].min() - t0) / 1e6:.2f} median {np.median(l[:nl, 1] - t0) / 1e6:.2f} max {(l[:nl, 1].max() - t0) / 1e6:.2f}")
    for cname, cost in (("attempts", att + (0.2 * len(taus) if "saveat" in kws else 0)), ("measured time per mode", tm)):
        # a cost in other units than attempts: the schedule's quotas are attempts, so scale each mode's cost to attempts when cutting (build_schedule takes the scale)
        s2 = sb.solvept(prob, bg, ks, cost=cost, **({"cost_attempts": att} if cname != "attempts" else {}), **kws); l, m = prof()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); s2 = sb.solvept(prob, bg, ks, cost=cost, **({"cost_attempts": att} if cname != "attempts" else {}), **kws); b.record(); torch.cuda.synchronize()
        l, m = prof()
        nl = int((l[:, 1] > 0).sum()); t0 = l[:nl, 0].min(); e = (l[:nl, 1] - t0) / 1e6; w = l[:nl, 2] / 1e6
        print(f"   static schedule, cost = {cname}: {a.elapsed_time(b):.2f} ms; list end times min {e.min():.2f} p10 {np.percentile(e, 10):.2f} median {np.median(e):.2f} p90 {np.percentile(e, 90):.2f} max {e.max():.2f} ms; "
              f"waiting for parked modes: total {w.sum():.1f} ms over {int((w > 0.01).sum())} lists, max {w.max():.2f} ms; ok {s2.success}")
