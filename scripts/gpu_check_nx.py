import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
from oracle import sbref
for nx in (8, 16):
    M = sb.ΛCDM(lmax=10, nx=nx)
    pars = sb.parameters_Planck18(M); pars["m_eV"] = 0.06
    t = time.time(); prob = sb.CosmologyProblem(M, pars); tb = time.time() - t
    bg = sb.solvebg(prob)
    ks = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time(); P, sol = sb.spectrum_matter(prob, ks, bgsol=bg, return_solution=True); torch.cuda.synchronize(); dt = time.time() - t
    obg = sbref.Background.from_knots(sbref.planck18(lmax=10, nx=nx, m_eV=0.06), bg.t, bg.y, bg.dy, bg.tau0, bg.kappa0)
    t = time.time(); Po, osol = sbref.spectrum_matter(obg, ks); to = time.time() - t
    rel = np.abs(P / Po - 1)
    print(f"nx={nx}: N={prob.N} nnz(W)={prob.info['nnz_full']} build {tb:.0f}s; GPU {dt:.3f}s ok={sol.success}; oracle {to:.1f}s; P(k) max rel {rel.max():.2e} median {np.median(rel):.2e}; steps equal for {(sol.stats[:,0]==osol['stats'][:,0]).mean()*100:.0f}% of modes")
