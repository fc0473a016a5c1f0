"""Robustness of the bit-identity between the two mappings: the config-4 sweep's cosmologies (Latin hypercube over h, Ω_c, Ω_b, A_s, n_s, w0, wa), including the
phantom-crossing ones whose modes fail (non-finite attempts, Unstable / DtLessThanMin return codes): split kernel against warp-per-mode kernel, bit for bit.  GPU box only."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import symboltz.jl_b200 as sb
from bench import config4_thetas
Mw = sb.w0waCDM(lmax=10); probw = sb.CosmologyProblem(Mw, sb.parameters_Planck18(Mw))
names, th = config4_thetas(512)
upd = sb.parameter_updater(probw, names)
ks = sb.loggrid(1e-4, 1.0, length=64) / sb.k0
f = lambda k: min(1e-2 / k, 1e-4)
nbad = ndiff = nfail_modes = 0
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    # the cosmologies whose sweeps fail first (w0 + wa crossing -1 inside the integration), then a sample of the regular ones
    cross = [i for i in range(len(th)) if (th[i, 5] + 1) * (th[i, 5] + th[i, 6] + 1) < 0 or th[i, 5] + th[i, 6] > -0.05]
    sel = cross[:40] + list(range(0, 512, 16))
    for i in sel:
        prob = upd(th[i])
        try:
            bg = sb.solvebg(prob)
        except Exception as e:
            nbad += 1
            continue
        a = sb.solvept(prob, bg, ks, ptivini=f, split=False, warn=False)
        b = sb.solvept(prob, bg, ks, ptivini=f, split=True, warn=False)
        same = np.array_equal(a.retcode, b.retcode) and np.array_equal(a.stats, b.stats) and np.array_equal(a.uend, b.uend, equal_nan=True)
        nfail_modes += int((a.retcode != 0).sum())
        if not same:
            ndiff += 1
            print("DIFF cosmology", i, th[i], "retcodes", a.retcode[a.retcode != b.retcode], b.retcode[a.retcode != b.retcode], flush=True)
print(f"{len(sel)} cosmologies ({len(cross[:40])} phantom-crossing candidates), {nbad} background failures, {nfail_modes} failing modes in total, {ndiff} cosmologies where the mappings differ")
