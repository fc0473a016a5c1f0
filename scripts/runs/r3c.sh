mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trbdf2" 2>&1 | tail -15
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import symboltz.jl_b200 as sb
M = sb.ΛCDM(lmax=10); prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M)); bg = sb.solvebg(prob)
ks = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
f = lambda k: min(1e-2 / k, 1e-4)
ref = sb.solvept(prob, bg, ks, ptivini=f, reltol=1e-9, abstol=1e-9)
for alg, tol in (("Rodas5P", 1e-5), ("TRBDF2", 1e-4), ("TRBDF2", 1e-5)):
    sb.solvept(prob, bg, ks, ptivini=f, reltol=tol, abstol=tol, alg=alg); torch.cuda.synchronize()
    t0 = time.perf_counter(); s = sb.solvept(prob, bg, ks, ptivini=f, reltol=tol, abstol=tol, alg=alg, split=False if alg == "Rodas5P" else None); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    err = (np.abs(s.uend - ref.uend).max(axis=1) / np.abs(ref.uend).max(axis=1)).max()
    print(f"config 1, {alg} tol {tol}: {1e3 * dt:.1f} ms, max attempts {int((s.stats[:, 0] + s.stats[:, 1]).max())}, state error vs tight Rodas5P {err:.1e}, ok {s.success}")
PY
