"""Reference robustness test (runtests.jl:507-546): Latin hypercube in a ±50 % box around the fiducial, ks = [1, 10, 100, 1000] H0/c,
every background and perturbation solve must succeed."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, warnings
import symboltz.jl_b200 as sb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for M in (sb.ΛCDM(lmax=10), sb.w0waCDM(lmax=10)):
    pars = sb.parameters_Planck18(M)
    prob = sb.CosmologyProblem(M, pars)
    names = ["h", "Omega_c", "Omega_b", "YHe", "Neff", "m_eV", "T0"] + (["w0", "wa"] if M.w0wa else [])
    fid = np.array([pars[k] for k in names])
    rng = np.random.default_rng(1)
    u = (rng.permuted(np.tile(np.arange(n), (len(names), 1)), axis=1).T + rng.random((n, len(names)))) / n
    th = fid * (0.5 + u) if not M.w0wa else np.where(np.arange(len(names)) < 7, fid * (0.5 + u), fid + 0.5 * np.abs(fid) * (2 * u - 1) * 0.5)
    ks = np.array([1.0, 10.0, 100.0, 1000.0])
    t = time.time()
    P, info = sb.spectrum_matter_sweep(prob, names, th, ks, return_info=True)
    print(M, f"{n} samples: {info}, finite {np.isfinite(P).all()}, {time.time()-t:.1f}s")
    bad = np.nonzero(~np.isfinite(P).all(axis=1))[0]
    for i in bad[:5]: print("  failed sample", dict(zip(names, th[i])))
