"""TEST INFRASTRUCTURE ONLY -- Python driver of the CPU oracle (oracle/sbref.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (symboltz.jl_b200) never does.

Restates, on the CPU, the reference's hot path (citations relative to /root/reference):
  solve / solvebg / solvept            src/solve.jl:380-402, 427-435, 496-569
  spectrum_primordial, spectrum_matter src/observables/fourier.jl:14-23, 74-101
  source_grid, source_kinterp          src/observables/fourier.jl:232-291, 419-467, 524-547
  SphericalBesselCache, los_integrate  src/observables/angular.jl:18-48, 109-185
  spectrum_cmb                         src/observables/angular.jl:198-223, 260-359
  grids                                src/utils.jl:269-292
Parity pinning: tests/test_oracle_golden.py checks this oracle against the reference's own golden
vectors (test/class_Pk.dat, test/class_Cl.dat) at the reference's tolerances (test/runtests.jl:872-886).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsbref.so")
NXMAX = 32
k0 = 1.0 / 2997.92458  # h/Mpc in units of H0/c (src/constants.jl:19)


def build(force=False):
    """Compile the C++ oracle with g++ (recipe also in oracle/Makefile)."""
    src = os.path.join(_HERE, "sbref.cpp")
    hdr = os.path.join(_HERE, "jet.hpp")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-o", _SO, src])
    return _SO


class SboParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("h", "Omega_c", "Omega_b", "T0", "YHe", "Neff", "m_eV", "Nh", "ln_As1e10", "ns", "w0", "wa", "cs2X")] + [
        ("lmax", C.c_int), ("nx", C.c_int), ("model", C.c_int), ("pad", C.c_int),
        ("x", C.c_double * NXMAX), ("W", C.c_double * NXMAX)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.sbo_nstate.restype = C.c_int
        _lib.sbo_background.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
def momentum_quadrature(N, L=100.0):
    """N-point Gauss rule for ∫dx x² f0(x) g(x), f0 = 1/(e^x+1), built in u = 1/(1+x/L)
    (reference src/models/neutrinos.jl:55-60; QuadGK.gauss with a weight function).
    Discretised Stieltjes procedure on a fine Gauss-Legendre grid + Golub-Welsch."""
    t, w = np.polynomial.legendre.leggauss(2000)
    u = 0.5 * (t + 1.0)
    wu = 0.5 * w
    x = L * (1 - u) / u
    with np.errstate(over="ignore"):
        f0 = 1.0 / (np.exp(x) + 1.0)
    wt = wu * (L / u**2) * x**2 * f0  # positive measure on u in (0,1)
    alpha = np.zeros(N)
    beta = np.zeros(N)
    pm1 = np.zeros_like(u)
    p0 = np.ones_like(u)
    for k in range(N):
        nrm = np.sum(wt * p0 * p0)
        alpha[k] = np.sum(wt * u * p0 * p0) / nrm
        if k > 0:
            beta[k] = nrm / np.sum(wt * pm1 * pm1)
        else:
            beta[0] = nrm
        p1 = (u - alpha[k]) * p0 - (beta[k] if k > 0 else 0.0) * pm1
        pm1, p0 = p0, p1
    J = np.diag(alpha) + np.diag(np.sqrt(beta[1:]), 1) + np.diag(np.sqrt(beta[1:]), -1)
    ev, evec = np.linalg.eigh(J)
    Ws = beta[0] * evec[0, :] ** 2
    xs = L * (1 - ev) / ev
    order = np.argsort(xs)
    return xs[order], Ws[order]


def planck18(**over):
    """Fiducial parameters, reference src/parameters.jl:3-19."""
    h = over.get("h", 0.6736)
    p = dict(h=h, Omega_c=0.1200 / h**2, Omega_b=0.0224 / h**2, T0=2.7255, YHe=0.2454, Neff=2.99, m_eV=0.02, Nh=3.0,
             ln_As1e10=float(np.log(2.099e-9 * 1e10)), ns=0.965, w0=-1.0, wa=0.0, cs2X=1.0, lmax=10, nx=4, model=0)
    p.update(over)
    return p


def make_params(d):
    p = SboParams()
    for k in ("h", "Omega_c", "Omega_b", "T0", "YHe", "Neff", "m_eV", "Nh", "ln_As1e10", "ns", "w0", "wa", "cs2X"):
        setattr(p, k, float(d[k]))
    p.lmax, p.nx, p.model, p.pad = int(d["lmax"]), int(d["nx"]), int(d["model"]), 0
    xs, Ws = momentum_quadrature(p.nx)
    for i in range(p.nx):
        p.x[i] = xs[i]
        p.W[i] = Ws[i]
    return p


class Background:
    """Result of the background solve: Hermite-spline knots of (a, _κ, XH⁺, XHe⁺, ΔT) (src/utils.jl:118-127)."""

    def __init__(self, pars, reltol=1e-7, abstol=1e-7, tini=1e-6, tmax=100.0):
        self.pars = dict(pars)
        self.p = make_params(pars)
        cap = 20000
        t = np.zeros(cap)
        y = np.zeros((cap, 5))
        dy = np.zeros((cap, 5))
        info = np.zeros(6)
        nb = lib().sbo_background(C.byref(self.p), C.c_double(tini), C.c_double(tmax), C.c_double(reltol), C.c_double(abstol), C.c_int(cap), _p(t), _p(y), _p(dy), _p(info))
        assert nb > 0
        self.t, self.y, self.dy = t[:nb].copy(), y[:nb].copy(), dy[:nb].copy()
        self.tau0, self.kappa0, self.taurec = info[0], info[1], info[2]
        self.retcode = int(info[3])
        self.naccept, self.nreject = int(info[4]), int(info[5])
        self.N = lib().sbo_nstate(C.byref(self.p))
        d = np.zeros(9)
        lib().sbo_derived(C.byref(self.p), _p(d))
        self.derived = dict(zip(("Omega_g", "Omega_nu", "Omega_h", "Omega_L", "fHe", "y0", "Irho0", "kpivot", "As"), d))

    @classmethod
    def from_knots(cls, pars, t, y, dy, tau0, kappa0, taurec=0.0):
        """Wrap externally computed spline knots (e.g. the product's background) so that the oracle's perturbation
        solver can be run on exactly the same background."""
        self = cls.__new__(cls)
        self.pars = dict(pars)
        self.p = make_params(pars)
        self.t, self.y, self.dy = np.ascontiguousarray(t), np.ascontiguousarray(y), np.ascontiguousarray(dy)
        self.tau0, self.kappa0, self.taurec, self.retcode, self.naccept, self.nreject = tau0, kappa0, taurec, 0, 0, 0
        self.N = lib().sbo_nstate(C.byref(self.p))
        d = np.zeros(9)
        lib().sbo_derived(C.byref(self.p), _p(d))
        self.derived = dict(zip(("Omega_g", "Omega_nu", "Omega_h", "Omega_L", "fHe", "y0", "Irho0", "kpivot", "As"), d))
        return self

    def observe(self, taus):
        taus = np.ascontiguousarray(taus, dtype=float)
        out = np.zeros((len(taus), 16))
        lib().sbo_bg_observe(C.byref(self.p), C.c_int(len(self.t)), _p(self.t), _p(self.y), _p(self.dy), C.c_double(self.kappa0), C.c_int(len(taus)), _p(taus), _p(out))
        names = ("a", "Hc", "Xe", "kappadot", "kappa", "v", "csb2", "Tb", "rho", "rho_h", "P_h", "fnu", "adot", "_kappa", "XH", "XHe")
        return {n: out[:, i] for i, n in enumerate(names)}

    def _spl(self):
        return (C.c_int(len(self.t)), _p(self.t), _p(self.y), _p(self.dy))


def solvept(bg, ks, ptivini=-np.inf, saveat=None, reltol=1e-5, abstol=1e-5, maxiters=100000, nthreads=0, alg="Rodas5P"):
    """Mirror of reference solvept (src/solve.jl:543-569). ptivini: number or callable k -> τini.
    Returns dict(uend[nk,N], usave[nk,nsave,N] or None, retcode[nk], stats[nk,4])."""
    ks = np.ascontiguousarray(ks, dtype=float)
    nk = len(ks)
    f = ptivini if callable(ptivini) else (lambda k: ptivini)
    tini = np.array([min(max(f(k), bg.t[0]), bg.t[-1]) for k in ks])  # clamp, src/solve.jl:527
    N = bg.N
    uend = np.zeros((nk, N))
    ret = np.zeros(nk, dtype=np.int32)
    stats = np.zeros((nk, 4), dtype=np.int64)
    if saveat is not None:
        saveat = np.ascontiguousarray(saveat, dtype=float)
        usave = np.zeros((nk, len(saveat), N))
        ns, ps, pu = len(saveat), _p(saveat), _p(usave)
    else:
        usave, ns, ps, pu = None, 0, None, None
    # alg: "Rodas5P" (reference default), "TRBDF2" / "KenCarp4" (reference ptalg(accuracy = 0 / 1), src/solve.jl:333-337; published schemes, unpinned against Julia)
    lib().sbo_solvept_alg(C.byref(bg.p), *bg._spl(), C.c_double(bg.tau0), C.c_double(bg.kappa0), C.c_int(nk), _p(ks), _p(tini), C.c_int(ns), ps,
                          C.c_double(reltol), C.c_double(abstol), C.c_long(maxiters), C.c_int(nthreads), pu, _p(uend), _p(ret), _p(stats), C.c_int({"Rodas5P": 0, "TRBDF2": 1, "KenCarp4": 2}[alg]))
    return dict(uend=uend, usave=usave, retcode=ret, stats=stats, tini=tini)


def spectrum_primordial(ks, bg):
    """P0(k) = 2π² As k⁻³ (k/kp)^(ns-1), src/observables/fourier.jl:14-23, src/models/inflation.jl:3-7."""
    ks = np.asarray(ks, dtype=float)
    return 2 * np.pi**2 * bg.derived["As"] / ks**3 * (ks / bg.derived["kpivot"]) ** (bg.pars["ns"] - 1)


def delta_m(bg, ks, u, tau=None):
    ks = np.ascontiguousarray(ks, dtype=float)
    u = np.ascontiguousarray(u)
    out = np.zeros(len(ks))
    lib().sbo_delta_m(C.byref(bg.p), *bg._spl(), C.c_double(bg.tau0 if tau is None else tau), C.c_int(len(ks)), _p(ks), _p(u), _p(out))
    return out


def spectrum_matter(bg, ks, kτini=1e-2, τinimax=1e-4, **kw):
    """P(k, τ0) of total matter (c+b+h), src/observables/fourier.jl:79-97. ks in H0/c; result in (c/H0)³."""
    sol = solvept(bg, ks, ptivini=lambda k: min(kτini / k, τinimax), **kw)
    P = spectrum_primordial(ks, bg) * delta_m(bg, ks, sol["uend"]) ** 2
    return P, sol


def sources(bg, ks, taus, usave):
    """CMB sources from saved states: returns array [nk, nτ, 7] = (ST, SE, SW, ISW, Doppler, polarization, Sψ)."""
    ks = np.ascontiguousarray(ks, dtype=float)
    taus = np.ascontiguousarray(taus, dtype=float)
    usave = np.ascontiguousarray(usave)
    out = np.zeros((len(ks), len(taus), 7))
    lib().sbo_sources(C.byref(bg.p), *bg._spl(), C.c_double(bg.tau0), C.c_double(bg.kappa0), C.c_double(bg.taurec), C.c_int(len(ks)), _p(ks), C.c_int(len(taus)), _p(taus), _p(usave), _p(out))
    return out


# --------------------------------------------------------------------------------------------- grids (src/utils.jl:269-292)
def lingrid(a, b, step=None, length=None):
    if step is not None:
        length = int(np.ceil((b - a) / step)) + 1
    return np.linspace(a, b, length)


def loggrid(a, b, **kw):
    x = np.exp(lingrid(np.log(a), np.log(b), **kw))
    x[0], x[-1] = a, b
    return x


def cospi(x):
    """cos(πx), argument reduced before multiplying by π (Julia's cospi, used by the reference's cosgrid, src/utils.jl:285): exact zeros at
    the half-integers."""
    x = np.abs(np.asarray(x, dtype=float))
    r = x - 2.0 * np.floor(x / 2.0)
    r = np.where(r > 1.0, 2.0 - r, r)
    sgn = np.where(r > 0.5, -1.0, 1.0)
    r = np.where(r > 0.5, 1.0 - r, r)
    return sgn * np.where(r <= 0.25, np.cos(np.pi * r), np.sin(np.pi * (0.5 - r)))


def cosgrid(a, b, length):
    return a + (b - a) * (1 - cospi(lingrid(0.0, 0.5, length=length)))


def chebpoints(order, a, b):
    """FastChebInterp.chebpoints: a + (b-a)(1+cos(πi/n))/2, i = 0..n (descending)."""
    i = np.arange(order + 1)
    return a + (b - a) * (1 + np.cos(np.pi * i / order)) / 2


def chebyshev_interp_matrix(ys_coarse, ys_fine):
    """Barycentric interpolation matrix B[nfine, ncoarse] (src/observables/fourier.jl:444-449, 524-535)."""
    n = len(ys_coarse) - 1
    w = np.array([1.0 if j % 2 == 0 else -1.0 for j in range(n + 1)])
    w[0] /= 2
    w[-1] /= 2
    B = np.zeros((len(ys_fine), n + 1))
    for i, y in enumerate(ys_fine):
        d = y - ys_coarse
        hit = np.where(d == 0)[0]
        if len(hit):
            B[i, hit[0]] = 1.0
        else:
            t = w / d
            B[i] = t / t.sum()
    return B


# --------------------------------------------------------------------------------------------- Bessel cache + LOS + C_l
class SphericalBesselCache:
    """src/observables/angular.jl:9-48 (uniform-x table of j_l and j_l', cubic Hermite evaluation)."""

    def __init__(self, ls, xmax=None, dx=2 * np.pi / 15, xcut=None, nthreads=None):
        """xcut: only tabulate x <= xcut (same grid points as the full table; saves memory, angular.jl:18-25).
        nthreads: the scipy ufunc releases the GIL, so the table is filled by a thread pool over blocks of x (default: all cores)."""
        from scipy.special import spherical_jn as _sj
        import concurrent.futures as cf
        import os
        nthreads = nthreads or len(os.sched_getaffinity(0))

        def spherical_jn(l, x):  # l: [nl, 1], x: [1, nx] -> [nl, nx], blocks of x in parallel
            xs_ = x[0]
            if nthreads <= 1 or len(xs_) < 4096:
                return _sj(l, x)
            cuts = np.linspace(0, len(xs_), 4 * nthreads + 1).astype(int)
            with cf.ThreadPoolExecutor(nthreads) as pool:
                parts = list(pool.map(lambda ab: _sj(l, xs_[None, ab[0]:ab[1]]), zip(cuts[:-1], cuts[1:])))
            return np.concatenate(parts, axis=1)
        self.l = np.asarray(ls)
        xmax = 20 * self.l[-1] if xmax is None else xmax
        n = int(xmax / dx)
        step = xmax / (n - 1)
        self.invdx = 1.0 / step
        self.dx = dx
        if xcut is None or xcut >= xmax:
            xs = np.append(np.linspace(0.0, xmax, n), xmax)
        else:
            m = int(np.ceil(xcut / step)) + 2
            xs = np.arange(m) * step
        self.x = xs
        L = self.l[:, None].astype(float)
        self.y = spherical_jn(self.l[:, None], xs[None, :])
        self.dy = L / (2 * L + 1) * spherical_jn(self.l[:, None] - 1, xs[None, :]) - (L + 1) / (2 * L + 1) * spherical_jn(self.l[:, None] + 1, xs[None, :])  # angular.jl:60

    def __call__(self, il, x):
        w = x * self.invdx
        i = np.trunc(w).astype(int)
        w = w - i
        wm1 = w - 1.0
        ym, yp, dm, dp = self.y[il, i], self.y[il, i + 1], self.dy[il, i], self.dy[il, i + 1]
        return (1 + 2 * w) * wm1 * wm1 * ym + w * w * (3 - 2 * w) * yp + w * wm1 * (wm1 * dm + w * dp) * self.dx


def los_integrate(Ss, ls, taus, ks, jl):
    """Θ_l(k) = Σ_τ w_τ S(τ,k) j_l(k(τ0-τ)); Ss[nτ, nk] → Is[nk, nl]  (src/observables/angular.jl:109-185)."""
    taus = np.asarray(taus)
    chis = taus[-1] - taus
    ws = np.empty_like(taus)
    ws[0] = 0.5 * (taus[1] - taus[0])
    ws[1:-1] = 0.5 * (taus[2:] - taus[:-2])
    ws[-1] = 0.5 * (taus[-1] - taus[-2])
    Is = np.zeros((len(ks), len(ls)))
    ils = np.arange(len(ls))[:, None]
    for ik, k in enumerate(ks):
        J = jl(ils, (k * chis)[None, :])  # [nl, nτ]
        Is[ik] = J @ (ws * Ss[:, ik])
    return Is


def los_limber(S, ls, taus, ks, l_limber):
    """Limber branch of los_integrate for l ≥ l_limber: Θ_l(k) = √(π/(2l+1)) S(τ0 − (l+½)/k, k)/k with the reference's cubic Hermite
    interpolation in χ (finite-difference slopes), src/observables/angular.jl:155-178.  S[nτ, nk] → [nk, nl] (zero for l < l_limber)."""
    taus = np.asarray(taus)
    chis = taus[-1] - taus
    nt = len(taus)
    out = np.zeros((len(ks), len(ls)))
    for ik, k in enumerate(ks):
        for il, l in enumerate(ls):
            if l < l_limber:
                continue
            chi = (l + 0.5) / k
            if chi > chis[0]:
                continue
            im = int(np.searchsorted(taus, taus[-1] - chi, side="left"))  # searchsortedfirst (0-based)
            Sm = S[im, ik]
            if im == 0:
                continue  # the reference leaves tmp[il] untouched (zero) in this branch
            ip = im - 1
            chim, chip, Sp = chis[im], chis[ip], S[ip, ik]
            dchi = chip - chim
            dSm = (S[im + 1, ik] - Sp) / (chis[im + 1] - chip) if im <= nt - 2 else (Sp - Sm) / dchi
            dSp = (Sm - S[im - 2, ik]) / (chim - chis[im - 2]) if ip >= 1 else (Sp - Sm) / dchi
            t = (chi - chim) / dchi
            t2, t3 = t * t, t * t * t
            Sv = (2 * t3 - 3 * t2 + 1) * Sm + (t3 - 2 * t2 + t) * dchi * dSm + (-2 * t3 + 3 * t2) * Sp + (t3 - t2) * dchi * dSp
            out[ik, il] = np.sqrt(np.pi / (2 * l + 1)) * Sv / k
    return out


def spectrum_cmb_lensing(modes, bg, jl, normalization="Cl", l_limber=10, order=130, kmin=1e-2, kmax=1e4, k0=2000.0, dkt0=np.pi, ntau=300, taucut=1e-2, nthreads=0, return_all=False):
    """spectrum_cmb with lensing modes (ψ): tanh-stretched Chebyshev grid to k = 1e4, Sψ source, Limber for l ≥ l_limber
    (src/observables/angular.jl:225-226, 267-269, 293-309).  modes: pairs of T, E, P (P = ψ)."""
    ls = np.asarray(jl.l)
    ks_fine = lingrid(kmin, kmax, step=dkt0 / bg.tau0)
    ts = bg.t[bg.t >= taucut]
    taus = ts[0] + (ts[-1] - ts[0]) * cosgrid(0.0, 1.0, length=ntau)
    taus[-1] = ts[-1]
    ys = chebpoints(order, np.tanh(kmin / k0), np.tanh(kmax / k0))
    kc = k0 * np.arctanh(ys)
    kc[-1], kc[0] = kmin, kmax
    sol = solvept(bg, kc, ptivini=-np.inf, saveat=taus, nthreads=nthreads)
    S = sources(bg, kc, taus, sol["usave"])
    Sc = [(kc[:, None] * S[:, :, 0]).T, (kc[:, None] ** 2 * S[:, :, 1]).T, S[:, :, 6].T.copy()]
    B = chebyshev_interp_matrix(ys, np.tanh(ks_fine / k0))
    Sf = []
    for a in Sc:
        a[-1, :] = 0.0
        Sf.append(a @ B.T)
    ThT = los_integrate(Sf[0], ls, taus, ks_fine, jl) / ks_fine[:, None]
    ThE = los_integrate(Sf[1], ls, taus, ks_fine, jl) * np.sqrt((ls + 2.0) * (ls + 1.0) * ls * (ls - 1.0))[None, :] / ks_fine[:, None] ** 2
    ThP = los_integrate(Sf[2], ls, taus, ks_fine, jl)
    lim = los_limber(Sf[2], ls, taus, ks_fine, l_limber)
    ThP[:, ls >= l_limber] = lim[:, ls >= l_limber]
    P0s = spectrum_primordial(ks_fine, bg)
    Th = {"T": ThT, "E": ThE, "P": ThP}
    out = np.stack([spectrum_cmb_from_theta(Th[m[0]], Th[m[1]], P0s, ls, ks_fine, normalization) for m in modes], axis=1)
    if return_all:
        return out, dict(ks_fine=ks_fine, kc=kc, taus=taus, Sf=Sf, Th=Th, sol=sol)
    return out


def natural_spline_integral(xs, ys):
    """∫ of the natural cubic spline through (xs, ys) (DataInterpolations CubicSpline + integral, angular.jl:212-213)."""
    from scipy.interpolate import CubicSpline
    return float(CubicSpline(xs, ys, bc_type="natural").integrate(xs[0], xs[-1]))


def spectrum_cmb_from_theta(ThA, ThB, P0s, ls, ks, normalization="Cl"):
    """C_l^AB = (2/π)∫dk k² P0 Θ_l^A Θ_l^B with a natural cubic spline through (0,0)+ks (angular.jl:198-223)."""
    ks0 = np.concatenate([[0.0], ks])
    Cl = np.zeros(len(ls))
    for il in range(len(ls)):
        f = np.concatenate([[0.0], 2 / np.pi * ks**2 * P0s * ThA[:, il] * ThB[:, il]])
        Cl[il] = natural_spline_integral(ks0, f)
    if normalization == "Dl":
        Cl = Cl * ls * (ls + 1) / (2 * np.pi)
    return Cl


def cmb_grids(bg, kmin=1e-2, kmax=2e3, dkt0=np.pi, ntau=300, taucut=1e-2):
    """Fine k-grid and τ-grid of spectrum_cmb (src/observables/angular.jl:275-290)."""
    ks_fine = lingrid(kmin, kmax, step=dkt0 / bg.tau0)
    ts = bg.t[bg.t >= taucut]
    taus = ts[0] + (ts[-1] - ts[0]) * cosgrid(0.0, 1.0, length=ntau)
    taus[-1] = ts[-1]
    return ks_fine, taus


def spectrum_cmb(modes, bg, jl, normalization="Cl", kmin=1e-2, kmax=2e3, order=60, dkt0=np.pi, ntau=300, taucut=1e-2,
                 direct=False, reltol=1e-5, abstol=1e-5, nthreads=0, return_all=False, ks=None, taus=None):
    """Mirror of spectrum_cmb(modes, prob, jl) for T/E modes (src/observables/angular.jl:260-341).
    direct=True solves every fine k instead of interpolating from the Chebyshev nodes.
    ks: solve exactly these (ascending) wavenumbers directly and use them as the k-quadrature grid (bounded samples of the
    direct workload for the CPU baseline timing in bench.py).  taus: use these line-of-sight times instead of the default grid."""
    ls = np.asarray(jl.l)
    tau0 = bg.tau0
    ks_fine, taus_default = cmb_grids(bg, kmin, kmax, dkt0, ntau, taucut)
    if taus is None:
        taus = taus_default
    if ks is not None:
        ks_fine, direct = np.ascontiguousarray(ks, dtype=float), True
    if taus is not None:  # line-of-sight times given by the caller (the default grid is anchored on a background SOLVER STEP, angular.jl:280-286)
        taus = np.ascontiguousarray(taus, dtype=float)
    ks_coarse = ks_fine if direct else chebpoints(order, kmin, kmax)
    sol = solvept(bg, ks_coarse, ptivini=-np.inf, saveat=taus, reltol=reltol, abstol=abstol, nthreads=nthreads)
    S = sources(bg, ks_coarse, taus, sol["usave"])  # [nk, nτ, 6]
    ST = (ks_coarse[:, None] * S[:, :, 0]).T        # k·ST, [nτ, nk]   (angular.jl:293)
    SE = (ks_coarse[:, None] ** 2 * S[:, :, 1]).T   # k²·SE
    if not direct:
        B = chebyshev_interp_matrix(ks_coarse, ks_fine)
        ST[-1, :] = 0.0
        SE[-1, :] = 0.0  # SE is Inf at χ = 0; the reference zeroes this row after interpolation (angular.jl:296)
        ST, SE = ST @ B.T, SE @ B.T
    ST[-1, :] = 0.0
    SE[-1, :] = 0.0  # angular.jl:296
    ThT = los_integrate(ST, ls, taus, ks_fine, jl) / ks_fine[:, None]
    ThE = los_integrate(SE, ls, taus, ks_fine, jl) * np.sqrt((ls + 2.0) * (ls + 1.0) * ls * (ls - 1.0))[None, :] / ks_fine[:, None] ** 2
    P0s = spectrum_primordial(ks_fine, bg)
    Th = {"T": ThT, "E": ThE}
    out = np.stack([spectrum_cmb_from_theta(Th[m[0]], Th[m[1]], P0s, ls, ks_fine, normalization) for m in modes], axis=1)
    if return_all:
        return out, dict(ks_fine=ks_fine, ks_coarse=ks_coarse, taus=taus, ST=ST, SE=SE, ThT=ThT, ThE=ThE, sol=sol, S=S)
    return out


def spline_ls(spectra_coarse, ls_coarse, ls_fine):
    """coarse-l → all-l cubic spline (angular.jl:348-359; DataInterpolations CubicSpline = natural)."""
    from scipy.interpolate import CubicSpline
    return np.stack([CubicSpline(np.asarray(ls_coarse, float), spectra_coarse[:, i], bc_type="natural")(np.asarray(ls_fine, float)) for i in range(spectra_coarse.shape[1])], axis=1)
