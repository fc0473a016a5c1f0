"""Integrator timing of the prebuilt model structures (BASELINE configs 1-4 + the reference's high-lmax test): P(k) grid of 100 modes and a 2019-mode batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
ks100 = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
ks2k = np.linspace(0.01, 2000, 2019)
f = lambda k: min(1e-2 / k, 1e-4)
for name, M, extra in (("LCDM l10 nx4", sb.ΛCDM(lmax=10), {}), ("w0wa l10 nx4", sb.w0waCDM(lmax=10), {}), ("LCDM l10 nx8 m=0.06", sb.ΛCDM(lmax=10, nx=8), {"m_eV": 0.06}),
                       ("LCDM l32 nx4", sb.ΛCDM(lmax=32), {}), ("LCDM l5 nx4", sb.ΛCDM(lmax=5), {})):
    pars = sb.parameters_Planck18(M); pars.update(extra)
    prob = sb.CosmologyProblem(M, pars); bg = sb.solvebg(prob)
    def run(ks, ptiv):
        sb.solvept(prob, bg, ks, ptivini=ptiv); torch.cuda.synchronize(); t = time.time(); s = sb.solvept(prob, bg, ks, ptivini=ptiv); torch.cuda.synchronize(); return s, time.time() - t
    s1, t1 = run(ks100, f); s2, t2 = run(ks2k, -np.inf)
    a1 = (s1.stats[:, 0] + s1.stats[:, 1]); a2 = (s2.stats[:, 0] + s2.stats[:, 1])
    print(f"{name:22s} N={prob.N:3d}: 100 modes {t1*1e3:7.1f} ms ({t1/a1.max()*1e6:5.1f} us/step isolated) | 2019 modes {t2*1e3:7.1f} ms = {2019/t2:7.0f} modes/s, {a2.sum()/t2/1e6:5.2f} M steps/s ok {s1.success and s2.success}")
