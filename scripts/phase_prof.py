"""Per-phase cycle breakdown of one mode (clock64-instrumented build of the integrator; diagnostic only)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
prob.lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsbm_prof.so"))
for k in (1000.0,):
    sol = sb.solvept(prob, bg, [k], trace=8)
    sol = sb.solvept(prob, bg, [k], trace=8)
    pr = sol.trace.ravel()[:9]; st = sol.stats[0]
    natt = st[0] + st[1]
    names = ["U/R combination", "f-eval (hub dots + ELL rows)", "rhs assemble", "B-solve", "Woodbury dots + update", "basis x5 (table lookup)", "factor + Z solves + capacitance", "8 stages total", "accept: norm, controller, basis+f0+dT"]
    print(f"k={k}: {st[0]} accepted, {st[1]} rejected steps")
    for n, v in zip(names, pr): print(f"  {n:42s} {v/natt:9.0f} cycles/step" + (f"  ({v/natt/8:6.0f} per stage)" if n in names[:5] else ""))
    print(f"  sum of step-level phases {(pr[5]+pr[6]+pr[7]+pr[8])/natt:9.0f} cycles/step")
