"""Pins the CPU oracle against the reference's own golden vectors and known-answer tests.
All CPU; file:line citations refer to /root/reference/test/runtests.jl."""
import ctypes as C
import os

import numpy as np
import pytest
from scipy.special import factorial, spherical_jn, zeta

from conftest import GOLDEN


def test_rodas5p_order_and_stiffness(oracle):
    """Re-verifies the recalled Rodas5P tableau: order 5, no order collapse on a stiff problem, 4th-order dense output."""
    L = oracle.lib()
    for lam in (1.0, 1e6):
        errs = []
        for n in (20, 40, 80, 160):
            y = np.zeros(2)
            L.sbo_rodas_fixed(C.c_double(lam), C.c_double(2.0), C.c_int(n), oracle._p(y), None)
            errs.append(np.hypot(y[0] - np.cos(2.0), y[1] - np.sin(4.0)))
        orders = [np.log2(errs[i] / errs[i + 1]) for i in range(3)]
        assert min(orders) > 4.7, (lam, errs, orders)
    derr = []
    for n in (10, 20, 40):
        y, mid = np.zeros(2), np.zeros(2)
        L.sbo_rodas_fixed(C.c_double(1.0), C.c_double(2.0), C.c_int(n), oracle._p(y), oracle._p(mid))
        tm = 2.0 - 1.0 / n
        derr.append(np.hypot(mid[0] - np.cos(tm), mid[1] - np.sin(2 * tm)))
    assert np.log2(derr[0] / derr[1]) > 3.7 and np.log2(derr[1] / derr[2]) > 3.7, derr


def test_momentum_quadrature_moments(oracle):
    """runtests.jl:549-559: Σ W x^(n-2) ≈ n!(1-2^-n)ζ(n+1) for the 4-point rule."""
    xs, Ws = oracle.momentum_quadrature(4)
    for n in range(2, 9):
        num = np.sum(Ws * xs ** (n - 2))
        anal = factorial(n) * (1 - 1 / 2**n) * zeta(n + 1)
        assert abs(num / anal - 1) < 10.0 ** (-6 + n - 1)


def test_time_and_optical_depth_today(obg5):
    """runtests.jl:332-340 and 153-160: a(τ0) ≈ 1 and not less; κ(τ0) = 0; no duplicated last knot."""
    o = obg5.observe(np.array([obg5.tau0]))
    assert abs(o["a"][0] - 1) < 1e-10 and o["a"][0] >= 1.0
    assert o["kappa"][0] == 0.0
    assert obg5.t[-1] == obg5.tau0 and obg5.t[-2] != obg5.t[-1]
    assert obg5.retcode == 0
    assert 3.0 < obg5.tau0 < 3.4  # ≈ 14 Gpc in units of c/H0


def test_initial_conditions_invariants(oracle, obg5):
    """runtests.jl:232-278: a ≈ √Ωr τ, τ ≈ 1/ℋ, F_l ∝ k^l, Φ/Ψ = 1 + 2fν/5, adiabaticity across species."""
    tini = obg5.t[0]
    o = obg5.observe(np.array([tini]))
    Or = obg5.derived["Omega_g"] + obg5.derived["Omega_nu"] + obg5.derived["Omega_h"] / obg5.derived["Irho0"] * 7 * np.pi**4 / 120
    assert abs(o["a"][0] - np.sqrt(Or) * tini) < 1e-10
    assert abs(tini - 1 / o["Hc"][0]) < 1e-10
    L, N = 5, obg5.N
    us = []
    for k in (0.1 / oracle.k0 * 0.6736 ** 0, 1.0 / oracle.k0):
        u = np.zeros(N)
        oracle.lib().sbo_pt_initial(C.byref(obg5.p), *obg5._spl(), C.c_double(tini), C.c_double(k), oracle._p(u))
        us.append((k, u))
    (k1, u1), (k2, u2) = us
    F1, F2 = u1[5:5 + L + 1], u2[5:5 + L + 1]
    for l in range(4):
        assert np.isclose(F1[l] / F2[l], (k1 / k2) ** l, rtol=1e-6)
    fnu = o["fnu"][0]
    for k, u in us:
        Psi = 20 * 0.5 / (15 + 4 * fnu)
        assert np.isclose(u[0] / Psi, 1 + 2 / 5 * fnu, atol=1e-4)
        # δ/(1+w) equal for c, b, γ, ν
        vals = [u[1], u[3], u[5] / (4 / 3), u[5 + 2 * (L + 1)] / (4 / 3)]
        assert np.ptp(vals) < 1e-3 * abs(vals[0])


def test_matter_power_spectrum_vs_class(oracle, obg5):
    """runtests.jl:872-876: isapprox(Pks, Pks_class; rtol = 1e-3) (2-norm) on the 516 CLASS wavenumbers, lmax = 5."""
    d = np.loadtxt(os.path.join(GOLDEN, "class_Pk.dat"))
    ks, Pc = d[:, 0], d[:, 1]
    P, sol = oracle.spectrum_matter(obg5, ks)
    assert (sol["retcode"] == 0).all()
    assert np.linalg.norm(P - Pc) <= 1e-3 * np.linalg.norm(Pc)
    assert np.abs(P / Pc - 1)[5:].max() < 5e-3  # element-wise, beyond CLASS's spline ringing at its lowest k


def test_pk_self_convergence(oracle, obg5):
    """runtests.jl:624-630: default tolerances vs tight tolerances agree within 1e-3 at every k."""
    ks = 10 ** np.linspace(-1, 3.5, 24)
    P, _ = oracle.spectrum_matter(obg5, ks)
    bgt = oracle.Background(oracle.planck18(lmax=5), reltol=1e-10, abstol=1e-10)
    Pt, _ = oracle.spectrum_matter(bgt, ks, kτini=0.0, τinimax=0.0, reltol=1e-9, abstol=1e-9)
    assert np.abs(P / Pt - 1).max() < 1e-3


def test_success_checking_k0_fails(oracle, obg5):
    """runtests.jl:358-361: k = 0 (and NaN, :31) must fail, positive k succeeds."""
    sol = oracle.solvept(obg5, np.array([0.0, np.nan, 1.0]))
    assert list(sol["retcode"] != 0) == [True, True, False]


def test_bessel_cache(oracle):
    """runtests.jl:123-145: Hermite table within 1e-5 of j_l; exact at x = 0."""
    ls = np.arange(10, 101, 10)
    jl = oracle.SphericalBesselCache(ls)
    xs = np.arange(0.0, jl.x[-2], 0.37)
    il = np.arange(len(ls))[:, None]
    assert np.abs(jl(il, xs[None, :]) - spherical_jn(ls[:, None], xs[None, :])).max() < 1e-5
    assert abs(jl(0, 0.0)) < 1e-16


def test_cmb_spectra_vs_class(oracle, obg5):
    """runtests.jl:879-885: D_l TT and EE vs CLASS, rtol = 2e-3 (2-norm), 200 log-spaced l then spline to l = 2..2500."""
    d = np.loadtxt(os.path.join(GOLDEN, "class_Cl.dat"))
    lc, TT, EE = d[:, 0], d[:, 1], d[:, 2]
    ls = np.unique(np.round(np.exp(np.linspace(np.log(lc[0]), np.log(lc[-1]), 200))).astype(int))
    jl = oracle.SphericalBesselCache(ls, xcut=2e3 * obg5.tau0 * 1.001)
    Dl = oracle.spectrum_cmb(["TT", "EE"], obg5, jl, normalization="Dl")
    Dl_all = oracle.spline_ls(Dl, ls, lc)
    assert np.linalg.norm(Dl_all[:, 0] - TT) <= 2e-3 * np.linalg.norm(TT)
    assert np.linalg.norm(Dl_all[:, 1] - EE) <= 2e-3 * np.linalg.norm(EE)


def test_lensing_spectrum_vs_class(oracle, obg5):
    """runtests.jl:886: D_l^{φφ} vs CLASS, rtol = 2e-3 (2-norm); lensing path = Sψ source, tanh-stretched Chebyshev grid to k = 1e4, Limber for l ≥ 10."""
    d = np.loadtxt(os.path.join(GOLDEN, "class_Cl.dat"))
    lc, PP = d[:, 0], d[:, 3]
    ls = np.unique(np.round(np.exp(np.linspace(np.log(2), np.log(2500), 32))).astype(int))
    jl = oracle.SphericalBesselCache(ls, xcut=1e4 * obg5.tau0 * 1.001)
    Dl = oracle.spectrum_cmb_lensing(["PP"], obg5, jl, normalization="Dl")
    Dl_all = oracle.spline_ls(Dl, ls, lc)
    assert np.linalg.norm(Dl_all[:, 0] - PP) <= 2e-3 * np.linalg.norm(PP)
