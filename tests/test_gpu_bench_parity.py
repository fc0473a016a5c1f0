"""GPU parity tests (-m gpu) on the BENCH model -- ΛCDM(lmax = 10), nx = 4, N = 82 unknowns per mode -- against the CPU oracle
(VERDICT r1 "weak" 1a-1f): BASELINE config 1 (100 log-spaced k, P(k)), config 2 on the reference's default Chebyshev path
(61 nodes, src/observables/angular.jl:267-273) and on the bench's "direct" path (all ≈2020 fine modes), a tight-tolerance
convergence test of C_l (TT, EE, TE), and the control experiment behind the default-tolerance bound.

What can and cannot agree to 1e-4 at the DEFAULT tolerance (reltol = abstol = 1e-5): the truncation error of either solve is
≈2e-3 in C_l (`test_cl_tight_tolerance...` measures it: default vs tight tolerance of ONE implementation), so two implementations
of the same adaptive scheme agree only as far as they take the same steps.  The first steps from τ = 1e-6 have a roundoff-dominated
error estimate; the oracle run against ITSELF with every wavenumber moved by a few ulp (`control`) shows the same ≈1e-4 scatter.
The default-tolerance tests therefore assert (i) 2-norm agreement at the north-star 1e-4, (ii) element-wise agreement within
max(1e-4, 2 × the control's own scatter), and the tight-tolerance test asserts element-wise 1e-5 (a tenth of the north star).
The oracle uses ITS OWN background solve here (`obg10`), not the product's knots, except where stated; the
line-of-sight time grid is the product's on both sides (the reference anchors it on a background solver step, so it is a property
of the background solver's step sequence, not of the hot path -- the effect of that alone is recorded as `own_tau_grid_*`).
Measured values are written to gpurun_out/parity_r2.json (copied to profiles/parity_r2.json)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LS129 = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))  # paper/aa57450-25.tex:786
_measured = {}


def _record(key, **vals):
    _measured[key] = {k: (float(v) if np.ndim(v) == 0 else np.asarray(v).tolist()) for k, v in vals.items()}
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_r2.json"), "w") as f:
        json.dump(_measured, f, indent=1)


@pytest.fixture(scope="module")
def prob10(sb):
    M = sb.ΛCDM(lmax=10)
    return sb.CosmologyProblem(M, sb.parameters_Planck18(M))


@pytest.fixture(scope="module")
def bg10(sb, prob10):
    return sb.solvebg(prob10)


@pytest.fixture(scope="module")
def obg10(oracle):
    """The oracle's OWN background solve (complex-step Jacobian, its own Rodas5P), independent of the product's."""
    return oracle.Background(oracle.planck18(lmax=10))


@pytest.fixture(scope="module")
def jl129(sb, bg10):
    return sb.SphericalBesselCache(LS129, xcut=2e3 * bg10.tau0 * 1.001)


@pytest.fixture(scope="module")
def ojl129(oracle, bg10):
    return oracle.SphericalBesselCache(LS129, xcut=2e3 * bg10.tau0 * 1.001)  # scipy table (≈30 s)


def _dev(D, Do):
    """(element-wise max relative deviation of TT and EE, TE deviation relative to max|TE|, 2-norm deviations of the three)."""
    el = np.abs(D[:, :2] / Do[:, :2] - 1).max(axis=0)
    te = np.abs(D[:, 2] - Do[:, 2]).max() / np.abs(Do[:, 2]).max()
    nrm = [np.linalg.norm(D[:, i] - Do[:, i]) / np.linalg.norm(Do[:, i]) for i in range(3)]
    return el, te, nrm


def test_config1_pk_100_logspaced_k_vs_oracle(sb, oracle, prob10, bg10, obg10):
    """BASELINE config 1: ΛCDM default lmax = 10, P(k, z = 0) on loggrid(1e-4, 1; 100) h/Mpc (src/observables/fourier.jl:79-101;
    SURVEY §8d).  North-star tolerance 1e-4 element-wise, against the oracle on the product's knots AND on its own background."""
    ks = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
    P, sol = sb.spectrum_matter(prob10, ks, bgsol=bg10, return_solution=True)
    assert sol.success and len(P) == 100
    same = oracle.Background.from_knots(oracle.planck18(lmax=10), bg10.t, bg10.y, bg10.dy, bg10.tau0, bg10.kappa0)
    Ps, osol = oracle.spectrum_matter(same, ks)
    Po, _ = oracle.spectrum_matter(obg10, ks)
    r_same, r_own = np.abs(P / Ps - 1), np.abs(P / Po - 1)
    _record("config1_pk", max_rel_same_knots=r_same.max(), median_rel_same_knots=np.median(r_same), max_rel_own_background=r_own.max(),
            steps_gpu_max=int(sol.stats[:, 0].max()), steps_oracle_max=int(osol["stats"][:, 0].max()))
    assert r_same.max() < 1e-4 and np.median(r_same) < 1e-7
    assert r_own.max() < 1e-4
    assert abs(bg10.tau0 / obg10.tau0 - 1) < 1e-7


def test_config2_cl_default_chebyshev_path_vs_oracle(sb, oracle, prob10, bg10, obg10, jl129, ojl129):
    """BASELINE config 2 on the reference's default path: 61 Chebyshev nodes on [1e-2, 2e3], barycentric interpolation to the fine
    grid, LOS, C_l at the 129 multipoles (src/observables/angular.jl:260-341), D_l TT/EE/TE vs the oracle pipeline with its own
    background.  2-norm ≤ 1e-4; element-wise ≤ max(1e-4, 2 × control) with control = oracle vs oracle with k-nodes moved by ulps."""
    Dl = sb.spectrum_cmb(["TT", "EE", "TE"], prob10, jl129, normalization="Dl", bgsol=bg10)
    # Same line-of-sight times on both sides ("same k/l grids", north star): the reference anchors its 300-point τ-grid on the first
    # background SOLVER STEP after τ = 1e-2 (src/observables/angular.jl:280-286), so two background solvers that agree to 1e-8 still
    # produce different quadrature grids; with the oracle's own grid the 300-point trapezoid error (≈2e-3 at low l, recorded below as
    # `own_tau_grid_*`) is what one compares, not the perturbation path.
    taus = np.minimum(sb.cmb_grids(bg10)[1], obg10.t[-1])
    oDl = oracle.spectrum_cmb(["TT", "EE", "TE"], obg10, ojl129, normalization="Dl", taus=taus)
    cDl = oracle.spectrum_cmb(["TT", "EE", "TE"], obg10, ojl129, normalization="Dl", taus=taus, kmin=1e-2 * (1 + 4e-16), kmax=2e3 * (1 - 4e-16))
    gDl = oracle.spectrum_cmb(["TT", "EE", "TE"], obg10, ojl129, normalization="Dl")
    el, te, nrm = _dev(Dl, oDl)
    cel, cte, cnrm = _dev(cDl, oDl)
    gel, gte, gnrm = _dev(Dl, gDl)
    per_l = np.abs(Dl[:, :2] / oDl[:, :2] - 1)
    _record("config2_chebyshev_default_tol", elementwise_TT_EE=el, TE_of_max=te, norm2=nrm, control_elementwise_TT_EE=cel, control_TE_of_max=cte, control_norm2=cnrm,
            l_exceeding_1e4_TT=LS129[per_l[:, 0] > 1e-4], l_exceeding_1e4_EE=LS129[per_l[:, 1] > 1e-4],
            own_tau_grid_elementwise_TT_EE=gel, own_tau_grid_TE_of_max=gte, own_tau_grid_norm2=gnrm)
    assert max(nrm) <= 1e-4
    assert (el <= np.maximum(1e-4, 2 * cel)).all() and te <= max(1e-4, 2 * cte)


def test_cl_tight_tolerance_converges_to_oracle(sb, oracle, prob10, bg10, obg10, jl129, ojl129):
    """reltol = abstol = 1e-9 on both sides (default Chebyshev path, bench model): the two implementations converge to the same
    answer -- TT, EE element-wise ≤ 1e-5, TE ≤ 1e-5 of its maximum.  The same-knots oracle isolates the perturbation
    path; the own-background oracle adds the (1e-7-tolerance) background solve.  Also measures the truncation error of the default
    tolerance, which is what bounds default-tolerance agreement between any two implementations."""
    tight = dict(reltol=1e-9, abstol=1e-9)
    Dl = sb.spectrum_cmb(["TT", "EE", "TE"], prob10, jl129, normalization="Dl", bgsol=bg10, ptopts=tight)
    same = oracle.Background.from_knots(oracle.planck18(lmax=10), bg10.t, bg10.y, bg10.dy, bg10.tau0, bg10.kappa0)
    sDl = oracle.spectrum_cmb(["TT", "EE", "TE"], same, ojl129, normalization="Dl", **tight)
    oDl = oracle.spectrum_cmb(["TT", "EE", "TE"], obg10, ojl129, normalization="Dl", taus=np.minimum(sb.cmb_grids(bg10)[1], obg10.t[-1]), **tight)  # same LOS times, see above
    el, te, nrm = _dev(Dl, sDl)
    el2, te2, nrm2 = _dev(Dl, oDl)
    Dd = sb.spectrum_cmb(["TT", "EE", "TE"], prob10, jl129, normalization="Dl", bgsol=bg10)
    tel, tte, _ = _dev(Dd, Dl)
    _record("config2_chebyshev_tight_tol", elementwise_TT_EE_same_knots=el, TE_of_max_same_knots=te, elementwise_TT_EE_own_background=el2, TE_of_max_own_background=te2,
            truncation_error_of_default_tolerance_TT_EE=tel, truncation_error_of_default_tolerance_TE=tte)
    assert el.max() <= 1e-5 and te <= 1e-5
    assert el2.max() <= 1e-5 and te2 <= 1e-5
    assert tel.max() > 1e-4  # the default tolerance itself is far coarser than the north-star 1e-4: that is the point of the control tests


def test_config2_direct_bench_configuration_vs_oracle(sb, oracle, prob10, bg10, obg10, jl129, ojl129):
    """The BENCH configuration (bench.py: every one of the ≈2020 fine modes solved directly, plan path with the static schedule as
    bench.py runs it) against the oracle: (a) S(τ,k) of 128 evenly spaced modes vs the oracle's sources (own step sequences), relative
    to each mode's maximum; (b) C_l of the full direct solve vs the oracle's full direct solve (all modes on the host cores)."""
    plan = sb.CMBPlan(prob10, bg10, jl129, modes=("TT", "EE", "TE"), direct=True, normalization="Dl")
    plan.upload()
    plan.run()
    plan.learn_schedule()  # what bench.py times
    Dl = plan.run_e2e()
    assert (plan.d_ret.cpu().numpy() == 0).all()
    same = oracle.Background.from_knots(oracle.planck18(lmax=10), bg10.t, bg10.y, bg10.dy, bg10.tau0, bg10.kappa0)
    oDl, ex = oracle.spectrum_cmb(["TT", "EE", "TE"], same, ojl129, normalization="Dl", direct=True, return_all=True)
    assert np.array_equal(ex["ks_fine"], plan.ks_fine) and np.array_equal(ex["taus"], plan.taus)  # grid indexing is bit-exact
    idx = np.linspace(0, plan.nk - 1, 128).round().astype(int)
    S = plan.d_S.cpu().numpy()[idx]                   # [128][2][nt] = (k·ST, k²·SE)
    oS = np.stack([ex["ST"][:, idx].T, ex["SE"][:, idx].T], axis=1)
    sdev = np.abs(S[:, :, :-1] - oS[:, :, :-1]).max(axis=2) / np.abs(oS[:, :, :-1]).max(axis=2)  # last τ: χ = 0 row (zeroed by the LOS)
    el, te, nrm = _dev(Dl, oDl)
    cDl = oracle.spectrum_cmb(["TT", "EE", "TE"], same, ojl129, normalization="Dl", direct=True, kmin=1e-2 * (1 + 4e-16), kmax=2e3 * (1 - 4e-16))
    cel, cte, cnrm = _dev(cDl, oDl)
    _record("config2_direct_bench", source_dev_of_mode_max_ST=[float(np.median(sdev[:, 0])), float(sdev[:, 0].max())], source_dev_of_mode_max_SE=[float(np.median(sdev[:, 1])), float(sdev[:, 1].max())],
            elementwise_TT_EE=el, TE_of_max=te, norm2=nrm, control_elementwise_TT_EE=cel, control_TE_of_max=cte, control_norm2=cnrm)
    assert np.median(sdev) < 1e-4 and sdev.max() < 5e-3  # per-mode step sequences differ (see module docstring); the C_l below is the criterion
    assert max(nrm) <= 1e-4
    assert (el <= np.maximum(1e-4, 2 * cel)).all() and te <= max(1e-4, 2 * cte)
