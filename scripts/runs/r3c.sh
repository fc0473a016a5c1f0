mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trbdf2 or kencarp4" 2>&1 | tail -12
python - > gpurun_out/sdirk_config1_r3c.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
f = lambda k: min(1e-2 / k, 1e-4)
ref = sb.solvept(prob, bg, ks, ptivini=f, reltol=1e-10, abstol=1e-10)
Pr = sb.spectrum_matter(prob, ks, bgsol=bg, reltol=1e-10, abstol=1e-10)
for alg, tol in (("Rodas5P", 1e-5), ("Rodas5P", 1e-4), ("KenCarp4", 1e-5), ("KenCarp4", 1e-4), ("TRBDF2", 1e-5), ("TRBDF2", 1e-4)):
    kw = dict(ptivini=f, reltol=tol, abstol=tol, alg=alg)
    sb.solvept(prob, bg, ks, **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); s = sb.solvept(prob, bg, ks, **kw); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    P = sb.spectrum_matter(prob, ks, bgsol=bg, ptopts=dict(alg=alg), reltol=tol, abstol=tol) if False else None
    err = (np.abs(s.uend - ref.uend).max(axis=1) / np.abs(ref.uend).max(axis=1)).max()
    print(f"config 1 (100 modes), {alg:8s} reltol = abstol = {tol:g}: {1e3 * dt:6.1f} ms, slowest mode {int((s.stats[:, 0] + s.stats[:, 1]).max()):5d} attempts, {int(s.stats[:, 3].max()):6d} linear solves; "
          f"max state error vs converged {err:.1e}; ok {s.success}", flush=True)
PY
cat gpurun_out/sdirk_config1_r3c.txt
