"""Per-mode attempted-step counts on the C_l grid (input to the static load-balancing cost model)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10)
out = {}
for tag, mod in (("planck", {}), ("var", {"Omega_c": 1.10, "h": 0.95, "Omega_b": 1.05})):
    pars = sb.parameters_Planck18(M)
    for k_, f in mod.items(): pars[k_] *= f
    prob = sb.CosmologyProblem(M, pars); bg = sb.solvebg(prob)
    ks, taus = sb.cmb_grids(bg)
    s = sb.solvept(prob, bg, ks, saveat=taus)
    out[tag + "_k"] = ks; out[tag + "_att"] = s.stats[:, 0] + s.stats[:, 1]; out[tag + "_tau0"] = bg.tau0
    print(tag, len(ks), out[tag + "_att"].sum(), out[tag + "_att"].max())
np.savez("gpurun_out/attempts.npz", **out)
