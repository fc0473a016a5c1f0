"""CPU tests of the product's host side: code generator vs the independent oracle, host background solver, grids,
interpolators, C-ABI loading.  No GPU compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_abi_exports_every_declared_symbol(sb, prob5):
    """Every function declared in include/symboltz_b200.h is exported by the built libraries."""
    hdr = open(os.path.join(ROOT, "include", "symboltz_b200.h")).read()
    names = re.findall(r"\b(sb[mlc]_\w+)\s*\(", hdr)
    assert len(names) > 30 and any(n.startswith("sbc_") for n in names)
    los, comm = C.CDLL(sb.build.build_los()), C.CDLL(sb.build.build_comm())
    for n in set(names):
        lib = prob5.lib if n.startswith("sbm_") else (comm if n.startswith("sbc_") else los)
        assert hasattr(lib, n), f"{n} not exported"
    # ownership rules of the sharded paths (same in every host language): strided modes, contiguous fine-k slices
    assert [comm.sbc_owned_count(10, r, 4) for r in range(4)] == [3, 3, 2, 2] and comm.sbc_owned_index(2, 1, 4) == 9
    assert [(comm.sbc_slice_begin(2019, r, 8), comm.sbc_slice_end(2019, r, 8)) for r in (0, 7)] == [(0, 252), (1766, 2019)]
    inf = (C.c_int * 16)()
    prob5.lib.sbm_info(inf)
    assert inf[0] == 47 and inf[4] == 5 and inf[5] == 4  # N = 5 + (lmax+1)(3+nx), reference test fixture lmax = 5


def test_model_sizes_match_reference_counts(sb):
    """N = 5 + (lmax+1)(3+nx); nnz(W) = 409 at lmax = 10, 124 unknowns at lmax = 16 (talk notebook)."""
    p10 = sb.CosmologyProblem(sb.ΛCDM(lmax=10), sb.parameters_Planck18(sb.ΛCDM()))
    assert p10.N == 82 and p10.info["nnz_full"] == 409
    pw = sb.CosmologyProblem(sb.w0waCDM(lmax=10), sb.parameters_Planck18(sb.w0waCDM()))
    assert pw.N == 84


def test_gpu_path_fails_loudly_without_gpu(sb, prob5, bg5):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="GPU-only"):
        sb.solvept(prob5, bg5, [1.0])
    with pytest.raises(RuntimeError, match="GPU-only"):
        sb.SphericalBesselCache(np.array([2, 3]))


def test_derived_parameters_match_oracle(prob5, obg5):
    for k in ("Omega_g", "Omega_nu", "Omega_h", "Omega_L", "fHe", "y0", "Irho0", "kpivot", "As"):
        assert np.isclose(prob5.derived[k], obg5.derived[k], rtol=1e-12), k


def test_momentum_quadrature(sb, oracle):
    from scipy.special import factorial, zeta
    for n in (4, 8, 16):
        xs, Ws = sb.momentum_quadrature(n)
        xo, Wo = oracle.momentum_quadrature(n)
        assert np.allclose(xs, xo, rtol=1e-9) and np.allclose(Ws, Wo, rtol=1e-8)
    xs, Ws = sb.momentum_quadrature(4)
    for m in range(2, 7):
        assert abs(np.sum(Ws * xs ** (m - 2)) / (factorial(m) * (1 - 1 / 2**m) * zeta(m + 1)) - 1) < 10.0 ** (-6 + m - 1)


def test_generated_bg_rhs_and_jacobian_match_oracle(prob5, obg5, oracle):
    """sympy program (forward-mode on the statement list) vs hand-written C++ with complex-step Jacobian."""
    for i in (50, 300, 500, 700, 900):
        y = obg5.y[i].copy()
        g, J, kd, obs = np.zeros(5), np.zeros((5, 5)), np.zeros(3), np.zeros(6)
        prob5.lib.sbm_debug_bg(_ptr(prob5.P), _ptr(y), _ptr(g), _ptr(J), _ptr(kd), _ptr(obs))
        g2, J2 = np.zeros(5), np.zeros((5, 5))
        oracle.lib().sbo_bg_rhs(C.byref(obg5.p), _ptr(y), _ptr(g2), _ptr(J2))
        assert np.allclose(g[:4], g2[:4], rtol=1e-7, atol=1e-300)  # RECFAST rates near Saha equilibrium are differences of large terms
        assert np.isclose(g[4], g2[4], rtol=1e-3)  # D(ΔT): difference of large terms at early times (Compton coupling), roundoff-limited
        m = J2 != 0
        assert np.allclose(J[m], J2[m], rtol=1e-5, atol=1e-6 * np.abs(J2).max())


def test_generated_f_J_dT_ic_match_oracle(sb, oracle):
    """The generated tables (J = J_local + p φᵀ + q ψᵀ, basis functions, ICs, Δm) against the oracle's hand-typed RHS,
    probed Jacobian and complex-step time gradient, for ΛCDM and w0waCDM."""
    rng = np.random.default_rng(1)
    for M, kw in ((sb.ΛCDM(lmax=5), dict(lmax=5)), (sb.ΛCDM(lmax=10), dict(lmax=10)), (sb.w0waCDM(lmax=10), dict(lmax=10, model=1, w0=-0.9, wa=0.1, cs2X=1.0))):
        prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
        obg = oracle.Background(oracle.planck18(**kw))
        P = prob.P.copy()
        P[prob.iP_kappa0], P[prob.iP_tau0] = obg.kappa0, obg.tau0
        N = prob.N
        assert N == obg.N
        for tau, k in ((1e-3, 50.0), (0.05, 300.0), (1.0, 5.0), (3.0, 1000.0)):
            u = rng.standard_normal(N)
            fo, Jo, dTo = np.zeros(N), np.zeros((N, N)), np.zeros(N)
            oracle.lib().sbo_pt_fjt(C.byref(obg.p), *obg._spl(), C.c_double(tau), C.c_double(k), _ptr(u), _ptr(fo), _ptr(Jo), _ptr(dTo))
            yo, ypo = np.zeros(5), np.zeros(5)
            prob.lib.sbm_debug_spline(C.c_int(len(obg.t)), _ptr(obg.t), _ptr(obg.y), _ptr(obg.dy), C.c_double(tau), _ptr(yo), _ptr(ypo))
            f, J, dT = np.zeros(N), np.zeros((N, N)), np.zeros(N)
            prob.lib.sbm_debug_fjt(_ptr(P), _ptr(yo), _ptr(ypo), C.c_double(tau), C.c_double(k), _ptr(u), _ptr(f), _ptr(J), _ptr(dT))
            assert np.abs(f - fo).max() <= 1e-11 * np.abs(fo).max()
            assert np.abs(J - Jo).max() <= 1e-11 * np.abs(Jo).max()
            assert np.abs(dT - dTo).max() <= 1e-9 * np.abs(dTo).max()
            assert ((J != 0) == (Jo != 0)).all()  # identical sparsity pattern
            u0, u0o = np.zeros(N), np.zeros(N)
            prob.lib.sbm_debug_initial(_ptr(P), _ptr(yo), C.c_double(tau), C.c_double(k), _ptr(u0))
            oracle.lib().sbo_pt_initial(C.byref(obg.p), *obg._spl(), C.c_double(tau), C.c_double(k), _ptr(u0o))
            assert np.abs(u0 - u0o).max() <= 1e-12 * np.abs(u0o).max()
            dm = C.c_double()
            prob.lib.sbm_debug_delta_m(_ptr(P), _ptr(yo), C.c_double(tau), C.c_double(k), _ptr(u), C.byref(dm))
            dmo = np.zeros(1)
            oracle.lib().sbo_delta_m(C.byref(obg.p), *obg._spl(), C.c_double(tau), C.c_int(1), _ptr(np.array([k])), _ptr(u), _ptr(dmo))
            assert np.isclose(dm.value, dmo[0], rtol=1e-11)


def test_woodbury_split_is_well_conditioned(sb, prob5, bg5):
    """W = B − pφᵀ − qψᵀ solved through B and a 2x2 capacitance matrix reproduces a dense solve to ~1e-12."""
    N = prob5.N
    rng = np.random.default_rng(0)
    for tau, k, h in ((1e-4, 1000.0, 1e-5), (0.05, 1000.0, 2e-3), (2.0, 1000.0, 2e-2), (3.0, 2000.0, 3e-3), (1.0, 10.0, 0.2)):
        y, yp = bg5.spline(tau)
        Jl, hubs = np.zeros((N, N)), np.zeros((4, N))
        prob5.lib.sbm_debug_split(_ptr(bg5.P), _ptr(y), _ptr(yp), C.c_double(tau), C.c_double(k), _ptr(Jl), _ptr(hubs))
        p, q, phi, psi = hubs
        B = np.eye(N) / (0.21193756319429014 * h) - Jl
        W = B - np.outer(p, phi) - np.outer(q, psi)
        r = rng.standard_normal(N)
        U, V = np.stack([p, q], 1), np.stack([phi, psi], 1)
        Z, yb = np.linalg.solve(B, U), np.linalg.solve(B, r)
        xw = yb + Z @ np.linalg.solve(np.eye(2) - V.T @ Z, V.T @ yb)
        xd = np.linalg.solve(W, r)  # dense pivoted LU; cond(W) reaches 1e12 in the tightly-coupled era, so compare solutions
        assert np.abs(xw - xd).max() <= 1e-8 * np.abs(xd).max()


def test_host_background_matches_oracle(bg5, obg5):
    """Host Rodas5P background (analytic generated Jacobian) vs the oracle (complex-step Jacobian): same τ0, a(τ), X_e within tolerance."""
    assert bg5.success and abs(bg5.tau0 / obg5.tau0 - 1) < 1e-8
    assert bg5.t[-1] == bg5.tau0 and bg5.y[-1, 0] >= 1.0 and abs(bg5.y[-1, 0] - 1) < 1e-10
    ts = np.exp(np.linspace(np.log(2e-6), np.log(0.999 * bg5.tau0), 400))
    Y = np.array([bg5.spline(t)[0] for t in ts])
    o = obg5.observe(ts)
    assert np.abs(Y[:, 0] / o["a"] - 1).max() < 1e-6          # runtests.jl:289 checkvar(a, 1e-6)
    assert np.abs(Y[:, 2] - o["XH"]).max() < 1e-5 and np.abs(Y[:, 3] - o["XHe"]).max() < 1e-5
    late = ts > 1e-2
    assert np.abs((Y[late, 1] - bg5.kappa0) - o["kappa"][late]).max() < 1e-4 * (1 + np.abs(o["kappa"][late]).max())


def test_grids(sb):
    """src/utils.jl:269-292"""
    g = sb.loggrid(1e-4, 1.0, length=100)
    assert g[0] == 1e-4 and g[-1] == 1.0 and len(g) == 100 and (np.diff(g) > 0).all()
    assert len(sb.lingrid(0.0, 1.0, step=0.3)) == 5 and sb.lingrid(0.0, 1.0, step=0.3)[-1] == 1.0
    c = sb.cosgrid(0.0, 1.0, length=300)
    assert c[0] == 0.0 and c[-1] == 1.0 and len(c) == 300 and (np.diff(c) > 0).all()  # cospi: the end points are exact (src/utils.jl:285)
    assert sb.cospi(0.5) == 0.0 and sb.cospi(1.0) == -1.0 and sb.cospi(1.5) == 0.0 and sb.cospi(2.0) == 1.0
    xs = np.linspace(0, 0.5, 1001)
    assert np.abs(sb.cospi(xs) - np.cos(np.pi * xs)).max() < 2e-16
    ch = sb.chebgrid(1.0, 3.0, order=4)
    assert np.allclose(ch, 2 + np.cos(np.pi * np.arange(4, -1, -1) / 4))
    with pytest.raises(ValueError):
        sb.lingrid(1.0, 0.0, length=3)


def test_chebyshev_interpolator_known_answers(sb):
    """runtests.jl:216-229: S = τ + k and τ·k through the Chebyshev interpolator (order 1 is exact for linear functions)."""
    ks = np.array([1.0, 10.0, 100.0])
    kint = sb.ChebyshevInterpolator(ks.min(), ks.max(), 1)
    assert kint.xs[0] == 100.0 and kint.xs[-1] == 1.0  # stored descending
    B = kint.matrix(ks)
    taus = np.array([1.0, 2.0])
    Sc = taus[:, None] + kint.xs[None, :]
    assert np.allclose(Sc @ B.T, taus[:, None] + ks[None, :])
    k60 = sb.ChebyshevInterpolator(1e-2, 2e3, 60)
    assert len(k60.xs) == 61 and np.isclose(k60.matrix(k60.xs), np.eye(61)).all()
    f = np.cos(k60.xs / 300.0)
    kf = np.linspace(1e-2, 2e3, 777)
    assert np.abs(k60.matrix(kf) @ f - np.cos(kf / 300.0)).max() < 1e-10


def test_natural_spline_weights_and_l_spline(sb):
    """The C_l k-integral is linear in the data: weights reproduce scipy's natural cubic spline integral / evaluation."""
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(3)
    x = np.concatenate([[0.0], np.sort(rng.random(40)) * 10 + 0.01])
    f = rng.standard_normal(len(x))
    w = sb.natural_spline_weights(x)
    assert np.isclose(w @ f, CubicSpline(x, f, bc_type="natural").integrate(x[0], x[-1]), rtol=1e-12)
    ls = np.array([2, 3, 5, 10, 20, 40, 80.0])
    y = np.stack([np.log(ls), ls**0.5], 1)
    lf = np.arange(2, 81)
    ref = np.stack([CubicSpline(ls, y[:, i], bc_type="natural")(lf) for i in range(2)], 1)
    assert np.allclose(sb.spline_ls(y, ls, lf), ref, rtol=1e-12)


def test_spectrum_primordial(sb, prob5):
    """src/observables/fourier.jl:14-23: P0 = 2π² As k⁻³ (k/kp)^(ns−1); kp = 0.05/Mpc in H0/c units."""
    k = np.array([1.0, 222.530031175772])
    P0 = sb.spectrum_primordial(k, prob5)
    assert np.isclose(P0[1], 2 * np.pi**2 * 2.099e-9 / k[1] ** 3, rtol=1e-9)
    assert np.isclose(prob5.derived["kpivot"], 0.05 / sb.k0 / 0.6736, rtol=1e-12)


def test_parameter_updater(sb, prob5):
    upd = sb.parameter_updater(prob5, ["Omega_c", "ns"])
    p2 = upd([0.3, 0.97])
    assert p2.pars["Omega_c"] == 0.3 and p2.pars["ns"] == 0.97 and p2.pars["h"] == prob5.pars["h"]
    assert p2.lib is prob5.lib  # the compiled engine is reused
    assert abs(p2.derived["Omega_L"] - (prob5.derived["Omega_L"] - (0.3 - prob5.pars["Omega_c"]))) < 1e-12


def test_adaptive_refinement_rule(sb):
    """refine_grid = the bisection rule of source_grid_adaptive (src/observables/fourier.jl:312-407) with the reference's
    known-answer sources S = (τ + k, τ·k) (runtests.jl:193-214): linear in k, so every initial interval is bisected exactly
    once; a curved source refines until linear interpolation meets isapprox; the 1024-point cap raises."""
    taus = np.array([1.0, 2.0])
    calls = []

    def lin(ks):
        calls.append(len(ks))
        return np.stack([taus[None, :] + ks[:, None], taus[None, :] * ks[:, None]], axis=-1)
    ks, S = sb.refine_grid(lin, np.linspace(1.0, 100.0, 3))
    assert np.array_equal(ks, [1.0, 25.75, 50.5, 75.25, 100.0]) and S.shape == (5, 2, 2) and calls == [3, 2]
    assert np.array_equal(S[:, :, 0], taus[None, :] + ks[:, None])
    # curved source, log transform (what spectrum_matter(prob, (kmin, kmax)) uses): leaf intervals satisfy the criterion, and
    # the grid is exactly the dyadic tree in ln k
    fn = lambda ks: (np.sin(3 * np.log(ks)) * ks**0.5)[:, None, None]
    ks, S = sb.refine_grid(fn, np.array([1.0, 10.0, 100.0]), atol=1e-3, rtol=1e-3, ktransform=(np.log, np.exp))
    assert (np.diff(ks) > 0).all() and ks[0] == 1.0 and ks[-1] == 100.0 and 17 < len(ks) < 1024
    y = np.log(ks) / np.log(100.0)
    assert np.allclose(y * 2**12, np.round(y * 2**12), atol=1e-6)  # dyadic in ln k
    mid = fn(np.exp((np.log(ks[:-1]) + np.log(ks[1:])) / 2))[:, 0, 0]
    lin2 = (S[:-1, 0, 0] + S[1:, 0, 0]) / 2
    assert (np.abs(mid - lin2) <= 4 * np.maximum(1e-3, 1e-3 * np.abs(mid))).all()  # children of accepted intervals are 4x closer (2nd order)
    with pytest.raises(RuntimeError, match="more than 1024"):
        sb.refine_grid(lambda ks: np.sin(1e4 * ks)[:, None, None], np.array([1.0, 10.0]), atol=1e-6)
    with pytest.raises(ValueError):
        sb.refine_grid(lin, np.array([1.0]))


def test_build_schedule_covers_every_mode_once_and_balances():
    """Host logic of the static preemptive schedule: every mode appears exactly once as a first piece / whole item, every split
    mode has exactly one continuation placed at the END of the previous list while its first piece is the FIRST item of the next
    one (so a continuation can never wait on a warp that has not started), and the simulated makespan is within 3 % of Σ/warps."""
    import symboltz.jl_b200 as sb
    rng = np.random.default_rng(0)
    k = np.linspace(0.01, 2000, 2019)
    cost = 160 + 2440 * (k / 2000) ** 0.55 + rng.normal(0, 5, len(k))
    W = 1184
    items, ibeg, T = sb.build_schedule(cost, W)
    first = items[items[:, 2] == 0]
    assert sorted(first[:, 0].tolist()) == list(range(len(k)))
    conts = items[items[:, 2] == 1]
    assert len(set(conts[:, 0].tolist())) == len(conts)
    assert set(conts[:, 0]) == set(first[first[:, 1] > 0, 0])
    fin = {}
    for w in range(W):
        lst = items[ibeg[w]:ibeg[w + 1]]
        assert (lst[:-1, 2] == 0).all()            # a continuation is only ever the last item
        assert (lst[1:, 1] == 0).all()             # a quota (first piece) only ever the first item
        if len(lst) and lst[0, 1] > 0:
            assert w > 0 and items[ibeg[w] - 1, 0] == lst[0, 0] and items[ibeg[w] - 1, 2] == 1
            fin[lst[0, 0]] = lst[0, 1]
    end = np.zeros(W)
    for w in range(W):
        t = 0.0
        for m, q, c in items[ibeg[w]:ibeg[w + 1]]:
            t = t + min(q, cost[m]) if (c == 0 and q > 0) else t + cost[m] if c == 0 else max(t, fin[m]) + cost[m] - fin[m]
        end[w] = t
    assert end.max() <= 1.03 * cost.sum() / W
    # degenerate inputs
    items, ibeg, _ = sb.build_schedule(np.ones(3), 8)
    assert len(items) == 3 and ibeg[-1] == 3
    m = sb.ModeCostModel(k, cost)
    assert np.abs(m(k) / cost - 1).max() < 0.2


def test_generated_path_layout_is_a_permutation_with_few_bank_collisions(sb):
    """The integrator's relabelled state order (codegen/lower.py): sb_nat is a permutation of the natural order, every elimination path is a
    contiguous index range inside [0, N), paths do not overlap, and the bank-aware tie-break leaves at most a handful of (path, path)
    pairs of the same half-warp whose ends are congruent modulo 16 doubles (each such pair costs one extra shared-memory wavefront per
    access).  With more paths than residues (nx = 8) or path lengths that are multiples of 16 (lmax = 32) collisions remain; only the layout is checked there."""
    import re
    for M in (sb.ΛCDM(lmax=10), sb.w0waCDM(lmax=10), sb.ΛCDM(lmax=5), sb.ΛCDM(lmax=10, nx=8), sb.ΛCDM(lmax=32)):
        prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
        so, info = sb.build.build_model(M.lmax, M.nx, M.w0wa)
        text = open(os.path.join(os.path.dirname(so), "sb_model_gen.h")).read()

        def arr(name):
            m = re.search(name + r"\[[^\]]*\]\s*=\s*\{([^}]*)\}", text)
            return [int(x.strip().rstrip("u"), 0) for x in m.group(1).split(",") if x.strip()]

        N = prob.N
        assert sorted(arr("sb_nat")) == list(range(N))
        bits8 = "#define SB_IDXBITS 8" in text
        heads = arr("sb_path_head")
        covered, ends, dummies = set(), [], []
        for slot, v in enumerate(heads):
            start, ln = (v & 255, (v >> 8) & 255) if bits8 else (v & 4095, (v >> 12) & 255)
            if ln:
                rng_ = set(range(start, start + ln))
                assert start + ln <= N and not (rng_ & covered)
                covered |= rng_
                ends.append(((slot % 32) // 16, (start + ln - 1) % 16, ln))  # 64-bit accesses are served per half-warp
            else:  # lane without a path: its unconditional loads (offsets down to −PL − 1 from start − 1) stay inside the arrays
                assert int(re.search(r"#define SB_PL (\d+)", text).group(1)) + 2 <= start < N
                dummies.append(((slot % 32) // 16, (start - 1) % 16))
        weight = sum(min(a[2], b[2]) for i, a in enumerate(ends) for b in ends[:i] if a[:2] == b[:2])
        if M.lmax <= 10 and M.nx == 4:  # the models of BASELINE configs 1, 2 and 4 (16-17 paths split over the two half-warps): conflict-free sweeps
            assert weight == 0, (M, weight, ends)
            assert {e[0] for e in ends} == {0, 1}
            assert not any((h, r) in {(e[0], e[1]) for e in ends} for (h, r) in set(d for d in dummies)), "a dummy lane shares a bank with a path"


def test_build_cli_and_freshness_stamps(sb, prob5, tmp_path):
    """`python -m symboltz.jl_b200.build` (what a non-Python host runs at CosmologyProblem time) prints the cached
    library; freshness is decided by a content hash of the sources, not by file times."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, "-m", "symboltz.jl_b200.build", "--lmax", "5"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    so = out.stdout.strip().splitlines()[0]
    assert os.path.exists(so) and so.endswith("libsbm_l5_x4_lcdm.so") and os.path.exists(so + ".stamp")
    src = tmp_path / "a.txt"
    tgt = tmp_path / "a.out"
    src.write_text("one")
    tgt.write_text("built")
    sb.build._stamp(str(tgt), [str(src)])
    os.utime(src, None)  # newer file time, same content: still fresh
    assert sb.build._newer(str(tgt), [str(src)])
    src.write_text("two")
    assert not sb.build._newer(str(tgt), [str(src)])


def test_lockstep_background_is_smooth_in_the_parameters(sb, prob5, bg5):
    """`solvebg_lock` (host; parameter lanes of BASELINE config 5): re-taking the primal's steps reproduces the primal to rounding, and
    the difference quotient of a moved parameter is linear in the step (no step-selection noise): δ = 1e-6 and 2e-6 give the same
    partials of all five background unknowns and of τ0 to 1e-4 of their size -- independent adaptive solves at reltol = 1e-7 could not
    resolve a 1e-6 step at all."""
    import math
    again = sb.solvebg_lock(prob5, bg5)
    assert len(again.t) == len(bg5.t) and abs(again.tau0 / bg5.tau0 - 1) < 1e-13
    assert np.abs(again.y - bg5.y).max() <= 1e-13 * np.abs(bg5.y).max()
    upd = sb.parameter_updater(prob5, ["Omega_c"])
    part = {}
    for d in (1e-6, 2e-6):
        lane = sb.solvebg_lock(upd([prob5.pars["Omega_c"] * math.exp(d)]), bg5)
        assert len(lane.t) == len(bg5.t) and np.array_equal(lane.t[:-1], bg5.t[:-1])  # same knots; only the event time moves
        part[d] = ((lane.y - bg5.y) / d, (lane.tau0 - bg5.tau0) / d)
    sc = np.abs(part[1e-6][0]).max(axis=0)
    assert (np.abs(part[1e-6][0] - part[2e-6][0]).max(axis=0) <= 1e-4 * sc).all()
    assert abs(part[1e-6][1] / part[2e-6][1] - 1) < 1e-5 and part[1e-6][1] < 0  # more matter -> earlier today
    with pytest.raises(RuntimeError):
        sb.solvebg_lock(upd([prob5.pars["Omega_c"] * 1.5]), bg5)  # a step that changes the number of solver steps is refused


def test_alternative_k_interpolators(sb):
    """the reference's other k-interpolators (src/observables/fourier.jl:199-231, 477-521, 549-571) as weight matrices
    for the same GPU k-interpolation / fused line-of-sight kernels.  Known answers: the natural spline reproduces scipy's, barycentric
    interpolation is exact for polynomials up to the order, the piecewise grid keeps the reference's node bookkeeping."""
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(2)
    xq = np.sort(rng.uniform(1.0, 9.0, 50))
    cs = sb.CubicSplineInterpolator(np.array([1.0, 1.5, 2.7, 4.0, 6.5, 9.0]), f=np.log)
    fv = rng.standard_normal(6)
    assert np.allclose(cs.matrix(xq) @ fv, CubicSpline(np.log(cs.xs), fv, bc_type="natural")(np.log(xq)), rtol=0, atol=1e-13)
    assert np.allclose(cs.matrix(cs.xs), np.eye(6), atol=1e-14) and (cs.minimum(), cs.maximum()) == (1.0, 9.0)
    assert len(sb.CubicSplineInterpolator(0.0, 1.0, n=4).xs) == 5
    with pytest.raises(ValueError):
        sb.CubicSplineInterpolator(np.array([2.0, 1.0]))
    eq = sb.EquispacedInterpolator(1.0, 9.0, 5)
    poly = lambda x: 3 - x + 0.2 * x**3 - 0.01 * x**5
    assert np.allclose(eq.matrix(xq) @ poly(eq.xs), poly(xq), rtol=1e-11) and np.array_equal(eq.matrix(eq.xs), np.eye(6))
    pw = sb.PiecewiseChebyshevInterpolator([1.0, 3.0, 9.0], [4, 6])
    assert len(pw.xs) == 5 + 7 - 1 and (np.diff(pw.xs) < 0).all() and pw.minimum() == 1.0 and pw.maximum() == 9.0
    assert pw.iranges[1] == slice(0, 7) and pw.iranges[0] == slice(6, 11) and pw.xs[6] == 3.0
    p4 = lambda x: 1 + x - 0.3 * x**2 + 0.05 * x**4
    assert np.allclose(pw.matrix(xq) @ p4(pw.xs), p4(xq), rtol=1e-11)
    assert np.array_equal(pw.matrix(np.array([3.0]))[0], np.eye(11)[6])
    with pytest.raises(ValueError):
        sb.PiecewiseChebyshevInterpolator([1.0, 2.0], [3, 3])


def test_background_differentiation(sb, prob5, bg5):
    """Reference "Background differentiation test" (test/runtests.jl:480-491): dτ0/dθ for θ = (h, Ω_c, Ω_b, T0, Neff, m_eV, YHe, ln 10¹⁰As, n_s)
    from the lockstep background lanes against central differences of independent tight-tolerance solves, atol 1e-2 (the reference's
    bound; measured ≈1e-6); parameters that do not enter the expansion (YHe, A_s, n_s) have exactly vanishing derivatives."""
    names = ["h", "Omega_c", "Omega_b", "T0", "Neff", "m_eV", "YHe", "ln_As1e10", "ns"]
    g = sb.sensitivity_background(prob5, names, bgsol=bg5)["tau0"]
    fd = np.zeros(len(names))
    for j, n in enumerate(names):
        x = prob5.pars[n]
        h = 1e-4 * abs(x)
        tp = sb.solvebg(sb.parameter_updater(prob5, [n])([x + h]), reltol=1e-11, abstol=1e-11).tau0
        tm = sb.solvebg(sb.parameter_updater(prob5, [n])([x - h]), reltol=1e-11, abstol=1e-11).tau0
        fd[j] = (tp - tm) / (2 * h)
    assert np.abs(g - fd).max() < 1e-2, (g, fd)
    assert np.abs(g[:6] / fd[:6] - 1).max() < 1e-3
    # A_s and n_s do not enter the background at all: exactly zero.  YHe enters only the recombination unknowns; the expansion a(τ) shares
    # the pivoted 5×5 solves with them, so τ0 moves in its last bits (1e-14 relative) and the quotient is rounding noise, not zero.
    assert (g[-2:] == 0).all() and abs(g[-3]) < 1e-6 and (np.abs(fd[-3:]) < 1e-2).all()


def test_oracle_trbdf2_converges_to_rodas5p(oracle):
    """The oracle's restatement of TRBDF2 (reference `ptalg(prob; accuracy = 0)`, src/solve.jl:333-335; OrdinaryDiffEq.jl absent: published scheme, unpinned
    against Julia) against its Rodas5P, which IS pinned to the reference's CLASS goldens: P(k) converges at the method's order (error ∝ tol^(2/3)) and the
    reference's own check for its alternative integrators -- success on k = 1, 10, 100, 1000 (test/runtests.jl:580-582) -- holds."""
    bg = oracle.Background(oracle.planck18(lmax=5))
    ks = np.array([1.0, 10.0, 100.0, 1000.0])
    Pt, _ = oracle.spectrum_matter(bg, ks, reltol=1e-9, abstol=1e-9)
    errs = []
    for tol in (1e-4, 1e-6):
        P, sol = oracle.spectrum_matter(bg, ks, reltol=tol, abstol=tol, alg="TRBDF2")
        assert (sol["retcode"] == 0).all()
        errs.append(np.abs(P / Pt - 1).max())
    assert errs[0] < 2e-2 and errs[1] < 1e-3 and errs[1] < 0.15 * errs[0], errs


def test_oracle_kencarp4_tableau_and_convergence(oracle):
    """KenCarp4 (reference `ptalg(prob; accuracy = 1)`, src/solve.jl:336-337; the algorithm of test/runtests.jl:580-590) restated from Kennedy & Carpenter's
    published ARK4(3)6L[2]SA tableau: consistency (row sums Σ_j a_ij + γ = c_i, Σ b = Σ b̂ = 1, stiffly accurate), order conditions up to 3 for b̂ and 4 for b
    on the quadrature conditions, and fourth-order convergence of P(k) to the CLASS-pinned Rodas5P result; success on k = 1, 10, 100, 1000."""
    import ctypes as C
    out = (C.c_double * 49)()
    oracle.lib().sbo_kencarp4_tableau(out)
    v = np.array(out)
    A, c, b, bh, g = v[:30].reshape(6, 5), v[30:36], v[36:42], v[42:48], v[48]
    Afull = np.zeros((6, 6)); Afull[:, :5] = A; Afull[np.arange(1, 6), np.arange(1, 6)] = g
    assert g == 0.25 and np.abs(Afull.sum(axis=1) - c).max() < 1e-15 and np.abs(Afull[5] - b).max() == 0
    assert abs(b.sum() - 1) < 1e-15 and abs(bh.sum() - 1) < 1e-15
    for q in (1, 2, 3):
        assert abs(b @ c**q - 1 / (q + 1)) < 1e-14      # quadrature conditions of order 4
    for q in (1, 2):
        assert abs(bh @ c**q - 1 / (q + 1)) < 1e-13     # ... and of order 3 for the companion
    assert abs(b @ (Afull @ c) - 1 / 6) < 1e-14 and abs(bh @ (Afull @ c) - 1 / 6) < 1e-13
    bg = oracle.Background(oracle.planck18(lmax=5))
    ks = np.array([1.0, 10.0, 100.0, 1000.0])
    Pt, _ = oracle.spectrum_matter(bg, ks, reltol=1e-10, abstol=1e-10)
    errs = []
    for tol in (1e-4, 1e-6):
        P, sol = oracle.spectrum_matter(bg, ks, reltol=tol, abstol=tol, alg="KenCarp4")
        assert (sol["retcode"] == 0).all()
        errs.append(np.abs(P / Pt - 1).max())
    assert errs[0] < 2e-3 and errs[1] < 2e-5 and errs[1] < 0.05 * errs[0], errs


def test_schedule_with_costs_in_other_units_and_ptalg(sb):
    """`build_schedule(cost, nlists, attempts=...)`: costs in any unit (e.g. measured time per mode); the quota of a split-off piece is converted to attempts with
    the mode's own attempts per unit of cost, so a cost vector scaled by a constant yields the same schedule.  `ptalg` mirrors src/solve.jl:326-341."""
    rng = np.random.default_rng(5)
    att = rng.integers(200, 3000, 500).astype(float)
    a_items, a_ibeg, _ = sb.build_schedule(att, 64)
    b_items, b_ibeg, _ = sb.build_schedule(att * 17.3e-6, 64, attempts=att)  # "seconds" instead of attempts
    assert np.array_equal(a_ibeg, b_ibeg) and np.array_equal(a_items[:, [0, 2]], b_items[:, [0, 2]])
    assert np.abs(a_items[:, 1] - b_items[:, 1]).max() <= 1  # quotas agree up to rounding
    for items, ibeg in ((a_items, a_ibeg), (b_items, b_ibeg)):
        first = items[items[:, 2] == 0]
        assert sorted(first[:, 0].tolist()) == list(range(500))
        assert ((first[:, 1] == 0) | (first[:, 1] < att[first[:, 0]])).all()
    assert [sb.ptalg(None, a) for a in (0, 1, 2, 3)] == ["TRBDF2", "KenCarp4", "Rodas5P", "Rodas5P"]
