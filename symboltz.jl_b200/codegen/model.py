"""Symbolic Einstein-Boltzmann model (sympy) -- the front end of the code generator.

This plays the role that ModelingToolkit/Symbolics play in the reference: it states the equations
of the ΛCDM / w0waCDM model family symbolically, so that lower.py can differentiate them and emit
CUDA C.  Equations follow the reference's component library (file:line relative to /root/reference):

  metric / gravity        src/models/metric.jl:21-26, src/models/gravity.jl:29-45
  generic species         src/models/generic_species.jl:33-60
  baryons + RECFAST       src/models/baryons.jl:14-116 (recombination), 128-138 (reionization), 145-212
  photons                 src/models/photons.jl:28-67
  massless neutrinos      src/models/neutrinos.jl:22-39
  massive neutrinos       src/models/neutrinos.jl:67-143
  dark energy             src/models/dark_energy.jl:6-17 (Λ), 24-67 (w0wa)
  assembly, ICs, sources  src/models/cosmologies.jl:50-106

Stage split (reference src/solve.jl:129-236): the 5 background unknowns y = (a, _κ, XH⁺, XHe⁺, ΔT)
are solved first on the host; in the perturbation system they are *inputs* evaluated from a cubic
Hermite spline (src/utils.jl:184-236), and every time derivative of a background quantity is expanded
through the background right-hand sides g(y) -- exactly what MTK's structural simplification produces.
"""
import sympy as sp

# ---------------------------------------------------------------- physical constants (src/constants.jl)
C = dict(
    c=299792458.0, h=6.62607015e-34, kB=1.380649e-23, GN=6.67430e-11, sigmaSB=5.670374419e-8,
    sigmaT=6.6524587321e-29, Mpc=3.0856775814913673e22, eV=1.602176634e-19, me=9.1093837015e-31,
    amu=1.66053906660e-27,
)
C["hbar"] = C["h"] / (2 * 3.141592653589793)
C["aR"] = 4 / C["c"] * C["sigmaSB"]
C["H100"] = 100 * 1e3 / C["Mpc"]
C["mH"] = 1.008 * C["amu"]
C["mHe"] = 4.0026022 * C["amu"]
C["k0"] = C["H100"] * C["Mpc"] / C["c"]


def _E(lam):
    return C["h"] * C["c"] / lam


TR = dict(
    EH_inf_2s=_E(91.17534e-9) - _E(121.56700e-9), EH_2s_1s=_E(121.56700e-9), lamH_2s_1s=121.56700e-9,
    EHe_inf_2s=_E(50.42590e-9) - _E(60.14045e-9), EHe_2s_1s=_E(60.14045e-9),
    lamHe_2p_1s=58.43344e-9, fHe_2p_1s=C["c"] / 58.43344e-9, EHe_2p_2s=_E(58.43344e-9) - _E(60.14045e-9),
    EHep_inf_1s=54.4178 * C["eV"],
    EHet_inf_2s=_E(260.0463e-9), lamHet_2p_1s=59.1411e-9, fHet_2p_1s=C["c"] / 59.1411e-9,
    EHet_2s_1s=_E(62.5563e-9), EHet_2p_2s=_E(59.1411e-9) - _E(62.5563e-9),
    LambdaH=8.2245809, LambdaHe=51.3, A2ps=1.798287e9, A2pt=177.58,
)

PI = sp.pi


def smoothifelse(x, v1, v2, k):
    """src/utils.jl:37"""
    return sp.Rational(1, 2) * ((v1 + v2) + (v2 - v1) * sp.tanh(k * x))


class Prog:
    """Straight-line symbolic program: ordered (symbol, expr) statements over inputs and earlier symbols."""

    def __init__(self, inputs):
        self.inputs = list(inputs)
        self.stmts = []  # (Symbol, expr)

    def let(self, name, expr):
        s = sp.Symbol("s_" + name, real=True)
        self.stmts.append((s, sp.sympify(expr)))
        return s


class Model:
    """Symbolic model. lmax: hierarchy cutoff, nx: massive-neutrino momentum bins, w0wa: CPL dark energy."""

    PAR_NAMES = ["h", "Omega_c", "Omega_b", "Omega_g", "Omega_nu", "Ch", "Omega_L", "T0", "YHe", "fHe", "y0",
                 "w0", "wa", "cs2X", "kappa0", "tau0"]

    def __init__(self, lmax=10, nx=4, w0wa=False):
        assert lmax >= 3
        self.lmax, self.nx, self.w0wa = lmax, nx, w0wa
        self.tau, self.k = sp.symbols("tau k", positive=True)
        self.y = sp.symbols("y_a y_kap y_XH y_XHe y_DT", real=True)
        self.par = {n: sp.Symbol("p_" + n, real=True) for n in self.PAR_NAMES}
        self.xs = [sp.Symbol(f"p_x{i}", positive=True) for i in range(nx)]
        self.Ws = [sp.Symbol(f"p_W{i}", positive=True) for i in range(nx)]
        self.dls = [sp.Symbol(f"p_dl{i}", real=True) for i in range(nx)]  # dlnf0/dlnx at x_i
        self.par_list = [self.par[n] for n in self.PAR_NAMES] + self.xs + self.Ws + self.dls
        self._background()
        self._perturbations()

    # ------------------------------------------------------------ background + thermodynamics
    def _background(self):
        """Builds the background chain as a straight-line symbolic program (self.prog): every `let`
        introduces a named temporary.  lower.py differentiates the program statement by statement
        (forward mode), which keeps code generation at seconds instead of differentiating one giant
        RECFAST expression tree."""
        p = self.par
        a, kap, XH, XHe, DT = self.y
        prog = self.prog = Prog(inputs=list(self.y) + [self.tau])
        let = prog.let
        pre = 3 / (8 * PI)
        a2 = let("a2", a * a)
        a3 = let("a3", a2 * a)
        a4 = let("a4", a2 * a2)
        rho_c = let("rc", pre * p["Omega_c"] / a3)
        rho_b = let("rb", pre * p["Omega_b"] / a3)
        rho_g = let("rg", pre * p["Omega_g"] / a4)
        rho_n = let("rn", pre * p["Omega_nu"] / a4)
        yh = let("yh", p["y0"] * a)
        E = [let(f"E{i}", sp.sqrt(x**2 + yh**2)) for i, x in enumerate(self.xs)]
        Irho = let("Irho", sum(W * e for W, e in zip(self.Ws, E)))
        IP = let("IP", sum(W * x**2 / e for W, x, e in zip(self.Ws, self.xs, E)))
        rho_h = let("rh", p["Ch"] * Irho / a4)
        P_h = let("Ph", p["Ch"] * IP / (3 * a4))
        if self.w0wa:
            w_X = let("wX", p["w0"] + p["wa"] * (1 - a))
            rho_X = let("rX", pre * p["Omega_L"] * a ** (-3 * (1 + p["w0"] + p["wa"])) * sp.exp(-3 * p["wa"] * (1 - a)))
        else:
            w_X = let("wX", sp.Integer(-1))
            rho_X = let("rX", pre * p["Omega_L"])
        rho = let("rho", rho_c + rho_b + rho_g + rho_n + rho_h + rho_X)
        adot = let("adot", sp.sqrt(8 * PI / 3 * rho) * a2)
        Hc = let("Hc", adot / a)
        wdX = let("wdX", -p["wa"] * adot if self.w0wa else sp.Integer(0))
        # thermodynamics
        H0SI = C["H100"] * p["h"]
        Tg = let("Tg", p["T0"] / a)
        DTg = let("DTg", -Tg * Hc)
        Tb = let("Tb", DT + Tg)
        nH = let("nH", (1 - p["YHe"]) * rho_b * H0SI**2 / C["GN"] / C["mH"])
        nHe = let("nHe", p["fHe"] * nH)
        beta = let("beta", 1 / (C["kB"] * Tb))
        lame = let("lame", C["h"] / sp.sqrt(2 * PI * C["me"] / beta))
        lame3 = let("lame3", lame**3)
        RHe = let("RHe", sp.exp(-beta * TR["EHep_inf_1s"]) / (nH * lame3))
        den = let("den", 1 + p["fHe"] + RHe)
        XHepp = let("XHepp", 2 * RHe * p["fHe"] / den / (1 + sp.sqrt(1 + 4 * RHe * p["fHe"] / den**2)))
        opz = let("opz", 1 / a)
        z1, dz, z2 = 7.6711, 0.5, 3.5
        Xre1 = let("Xre1", smoothifelse((1 + z1) ** 1.5 - sp.sqrt(opz) * opz, 0, 1 + p["fHe"], 1 / (1.5 * (1 + z1) ** 0.5 * dz)))
        Xre2 = let("Xre2", smoothifelse((1 + z2) - opz, 0, p["fHe"], 1 / dz))
        Xe = let("Xe", XH + p["fHe"] * XHe + XHepp + Xre1 + Xre2)
        ne = let("ne", Xe * nH)
        kapdot = let("kd", -a / H0SI * ne * C["sigmaT"] * C["c"])
        muc2 = let("muc2", C["mH"] * C["c"] ** 2 / (1 + (C["mH"] / C["mHe"] - 1) * p["YHe"] + Xe * (1 - p["YHe"])))
        DTb = let("DTb", -2 * Tb * Hc - a / p["h"] * (sp.Rational(8, 3) * C["sigmaT"] * C["aR"] / C["H100"]) * Tg**4 / (C["me"] * C["c"]) * Xe / (1 + p["fHe"] + Xe) * DT)
        csb2 = let("cs2", C["kB"] / muc2 * (Tb - DTb / (3 * Hc)))
        dDT = let("dDT", DTb - DTg)
        # RECFAST rate equations
        HSI = let("HSI", H0SI * Hc / a)
        Tr = let("Tr", Tb / 1e4)
        alphaH = let("alphaH", 1.125 * 1e-19 * 4.309 * Tr ** (-0.6166) / (1 + 0.6703 * Tr**0.5300))
        betaH = let("betaH", alphaH / lame3 * sp.exp(-beta * TR["EH_inf_2s"]))
        lna = let("lna", sp.log(a))
        KHfit = let("KHfit", 1 - 0.14 * sp.exp(-(((lna + 7.28) / 0.18) ** 2)) + 0.079 * sp.exp(-(((lna + 6.73) / 0.33) ** 2)))
        KH = let("KH", KHfit / (8 * PI) * TR["lamH_2s_1s"] ** 3 / HSI)
        CHfull = let("CHfull", (1 + KH * TR["LambdaH"] * nH * (1 - XH)) / (1 + KH * (TR["LambdaH"] + betaH) * nH * (1 - XH)))
        CH = let("CH", smoothifelse(XH - 0.99, CHfull, 1, 1e3))
        dXH = let("dXH", -a / H0SI * CH * (alphaH * XH * ne - betaH * (1 - XH) * sp.exp(-beta * TR["EH_2s_1s"])))
        sT2 = let("sT2", sp.sqrt(Tb / 3.0))
        sT1 = let("sT1", sp.sqrt(Tb / 10**5.114))

        def alphaHefit(q, pp):
            return q / (sT2 * (1 + sT2) ** (1 - pp) * (1 + sT1) ** (1 + pp))

        eps = 1e-9
        alphaHe = let("alphaHe", alphaHefit(10 ** (-16.744), 0.711))
        betaHe = let("betaHe", 4 * alphaHe / lame3 * sp.exp(-beta * TR["EHe_inf_2s"]))
        invKHe0 = let("invKHe0", 8 * PI * HSI / TR["lamHe_2p_1s"] ** 3)
        tauHe = let("tauHe", 3 * TR["A2ps"] * nHe * (1 - XHe + eps) / invKHe0)
        invKHe1 = let("invKHe1", -sp.exp(-tauHe) * invKHe0)
        gcom = let("gcom", 3 * p["fHe"] * (1 - XHe + eps) * C["c"] ** 2 / (8 * PI * sp.sqrt(2 * PI / (beta * C["mHe"] * C["c"] ** 2)) * (1 - XH + eps)))
        g2ps = let("g2ps", gcom * TR["A2ps"] / (1.436289e-22 * TR["fHe_2p_1s"] ** 3))
        invKHe2 = let("invKHe2", TR["A2ps"] / (1 + 0.36 * g2ps**0.86) * 3 * nHe * (1 - XHe))
        KHe = let("KHe", 1 / (invKHe0 + invKHe1 + invKHe2))
        e2p2s = let("e2p2s", sp.exp(-beta * TR["EHe_2p_2s"]))
        CHefull = let("CHefull", (e2p2s + KHe * TR["LambdaHe"] * nHe * (1 - XHe)) / (e2p2s + KHe * (TR["LambdaHe"] + betaHe) * nHe * (1 - XHe)))
        CHe = let("CHe", smoothifelse(XHe - 0.99, CHefull, 1, 1e3))
        DXHes = let("DXHes", -a / H0SI * CHe * (alphaHe * XHe * ne - betaHe * (1 - XHe) * sp.exp(-beta * TR["EHe_2s_1s"])))
        alphaHet = let("alphaHet", alphaHefit(10 ** (-16.306), 0.761))
        betaHet = let("betaHet", sp.Rational(4, 3) * alphaHet / lame3 * sp.exp(-beta * TR["EHet_inf_2s"]))
        tauHet = let("tauHet", TR["A2pt"] * nHe * (1 - XHe + eps) * 3 * TR["lamHet_2p_1s"] ** 3 / (8 * PI * HSI))
        pHet = let("pHet", (1 - sp.exp(-tauHet)) / tauHet)
        g2pt = let("g2pt", gcom * TR["A2pt"] / (1.484872e-22 * TR["fHet_2p_1s"] ** 3))
        CHetnum = let("CHetnum", TR["A2pt"] * (pHet + 1 / (1 + 0.66 * g2pt**0.9) / 3) * sp.exp(-beta * TR["EHet_2p_2s"]))
        CHet = let("CHet", (eps + CHetnum) / (eps + CHetnum + betaHet))
        DXHet = let("DXHet", -a / H0SI * CHet * (alphaHet * XHe * ne - betaHet * (1 - XHe) * 3 * sp.exp(-beta * TR["EHet_2s_1s"])))
        dXHe = let("dXHe", DXHes + DXHet)

        self.g = [adot, kapdot, dXH, dXHe, dDT]  # background RHS (program temporaries), autonomous in τ
        # named background observables used by the perturbation rows: name -> program temporary
        self.S = dict(a=a, Hc=Hc, kd=kapdot, cs2=csb2, rc=rho_c, rb=rho_b, rg=rho_g, rn=rho_n, rh=rho_h, Ph=P_h,
                      rX=rho_X, wX=w_X, wdX=wdX, Irho=Irho, IP=IP, Xe=Xe, Tb=Tb)
        for i, e in enumerate(E):
            self.S[f"E{i}"] = e

    # ------------------------------------------------------------ perturbations
    def _perturbations(self):
        L, nx, k, tau, p = self.lmax, self.nx, self.k, self.tau, self.par
        S = self.S  # named background temporaries of self.prog
        names = ["Phi", "dc", "tc", "db", "tb"] + [f"F{l}" for l in range(L + 1)] + [f"G{l}" for l in range(L + 1)] + [f"N{l}" for l in range(L + 1)]
        for i in range(nx):
            names += [f"psi{i}_{l}" for l in range(L + 1)]
        if self.w0wa:
            names += ["dX", "tX"]
        self.unames = names
        self.N = len(names)
        self.u = [sp.Symbol("u_" + n, real=True) for n in names]
        U = dict(zip(names, self.u))
        self.U = U
        a, Hc, kd = S["a"], S["Hc"], S["kd"]
        E = [S[f"E{i}"] for i in range(nx)]
        F = [U[f"F{l}"] for l in range(L + 1)]
        G = [U[f"G{l}"] for l in range(L + 1)]
        Nn = [U[f"N{l}"] for l in range(L + 1)]
        psi = [[U[f"psi{i}_{l}"] for l in range(L + 1)] for i in range(nx)]
        # total density perturbation and anisotropic stress (cosmologies.jl:92,94; neutrinos.jl:117-124)
        drho_h = p["Ch"] / a**4 * sum(self.Ws[i] * E[i] * psi[i][0] for i in range(nx))
        Pi_h = p["Ch"] / a**4 * sp.Rational(2, 3) * sum(self.Ws[i] * self.xs[i] ** 2 / E[i] * psi[i][2] for i in range(nx))
        drho = U["dc"] * S["rc"] + U["db"] * S["rb"] + F[0] * S["rg"] + Nn[0] * S["rn"] + drho_h
        if self.w0wa:
            drho += U["dX"] * S["rX"]
        Pi = sp.Rational(4, 3) * S["rg"] * F[2] / 2 + sp.Rational(4, 3) * S["rn"] * Nn[2] / 2 + Pi_h
        self.Psi_expr = U["Phi"] - 12 * PI * a**2 * Pi / k**2                      # gravity.jl:39
        Psi, Phd = sp.Symbol("hub_Psi", real=True), sp.Symbol("hub_Phd", real=True)  # "hub" observed variables
        self.hubs = (Phd, Psi)
        self.Phd_expr = -4 * PI / 3 * a**2 / Hc * drho - k**2 / (3 * Hc) * U["Phi"] - Hc * Psi  # gravity.jl:38
        f = {}
        f["Phi"] = Phd
        f["dc"] = -(U["tc"] - 3 * Phd)
        f["tc"] = -Hc * U["tc"] + k**2 * Psi
        thg = 3 * k * F[1] / 4
        f["db"] = -(U["tb"] - 3 * Phd) - 3 * Hc * S["cs2"] * U["db"]
        f["tb"] = -Hc * U["tb"] + S["cs2"] * k**2 * U["db"] + k**2 * Psi - kd * 4 * S["rg"] / (3 * S["rb"]) * (thg - U["tb"])
        Pig = F[2] + G[0] + G[2]
        self.Pig_expr = Pig
        f["F0"] = -k * F[1] + 4 * Phd
        f["F1"] = k / 3 * (F[0] - 2 * F[2] + 4 * Psi) - sp.Rational(4, 3) * kd / k * (U["tb"] - thg)
        for l in range(2, L):
            f[f"F{l}"] = k / (2 * l + 1) * (l * F[l - 1] - (l + 1) * F[l + 1]) + kd * (F[l] - (Pig / 10 if l == 2 else 0))
        f[f"F{L}"] = k * F[L - 1] - (L + 1) / tau * F[L] + kd * F[L]
        f["G0"] = -k * G[1] + kd * (G[0] - Pig / 2)
        f["G1"] = k / 3 * (G[0] - 2 * G[2]) + kd * G[1]
        for l in range(2, L):
            f[f"G{l}"] = k / (2 * l + 1) * (l * G[l - 1] - (l + 1) * G[l + 1]) + kd * (G[l] - (Pig / 10 if l == 2 else 0))
        f[f"G{L}"] = k * G[L - 1] - (L + 1) / tau * G[L] + kd * G[L]
        f["N0"] = -k * Nn[1] + 4 * Phd
        f["N1"] = k / 3 * (Nn[0] - 2 * Nn[2] + 4 * Psi)
        for l in range(2, L):
            f[f"N{l}"] = k / (2 * l + 1) * (l * Nn[l - 1] - (l + 1) * Nn[l + 1])
        f[f"N{L}"] = k * Nn[L - 1] - (L + 1) / tau * Nn[L]
        for i in range(nx):
            xE, Ex, dl = self.xs[i] / E[i], E[i] / self.xs[i], self.dls[i]
            f[f"psi{i}_0"] = -k * xE * psi[i][1] - Phd * dl
            f[f"psi{i}_1"] = k / 3 * xE * (psi[i][0] - 2 * psi[i][2]) - k / 3 * Ex * Psi * dl
            for l in range(2, L):
                f[f"psi{i}_{l}"] = k / (2 * l + 1) * xE * (l * psi[i][l - 1] - (l + 1) * psi[i][l + 1])
            f[f"psi{i}_{L}"] = k / (2 * L + 1) * xE * (L * psi[i][L - 1] - (L + 1) * ((2 * L + 1) * Ex * psi[i][L] / (k * tau) - psi[i][L - 1]))
        if self.w0wa:
            w, cs2 = S["wX"], p["cs2X"]
            ca2 = w - S["wdX"] / (3 * Hc * (1 + w))
            f["dX"] = -(1 + w) * (U["tX"] - 3 * Phd) - 3 * Hc * (cs2 - w) * U["dX"] - 9 * (Hc / k) ** 2 * (1 + w) * (cs2 - ca2) * U["tX"]
            f["tX"] = -Hc * (1 - 3 * cs2) * U["tX"] + cs2 / (1 + w) * k**2 * U["dX"] + k**2 * Psi
        self.f = [sp.expand(f[n]) for n in names]

        # ---------------- initial conditions (closed form given Ψ; see SURVEY.md App. C for reference lines)
        fnu = (S["rn"] + S["rh"]) / (S["rg"] + S["rn"] + S["rh"])
        Psi0 = 20 * sp.Rational(1, 2) / (15 + 4 * fnu)
        kt, kdk = k * tau, k / kd
        ic = {n: sp.Integer(0) for n in names}
        ic["dc"] = ic["db"] = -sp.Rational(3, 2) * Psi0
        ic["tc"] = ic["tb"] = sp.Rational(1, 2) * k**2 * tau * Psi0
        ic["F0"] = -2 * Psi0
        ic["F1"] = sp.Rational(2, 3) * kt * Psi0
        ic["F2"] = -sp.Rational(8, 15) * kdk * ic["F1"]
        ic["F3"] = -sp.Rational(3, 7) * kdk * ic["F2"]
        ic["G0"] = sp.Rational(5, 16) * ic["F2"]
        ic["G1"] = -sp.Rational(1, 16) * kdk * ic["F2"]
        ic["G2"] = sp.Rational(1, 16) * ic["F2"]
        ic["G3"] = -sp.Rational(3, 7) * kdk * ic["G2"]
        ic["N0"] = -2 * Psi0
        ic["N1"] = 4 * (sp.Rational(1, 2) * k**2 * tau * Psi0) / (3 * k)
        ic["N2"] = 2 * (kt**2 * Psi0 / 15)
        ic["N3"] = sp.Rational(3, 7) * kt * ic["N2"]
        for i in range(nx):
            dl = self.dls[i]
            ic[f"psi{i}_0"] = -sp.Rational(1, 4) * (-2 * Psi0) * dl
            ic[f"psi{i}_1"] = -sp.Rational(1, 3) * E[i] / self.xs[i] * (sp.Rational(1, 2) * kt * Psi0) * dl
            ic[f"psi{i}_2"] = -sp.Rational(1, 2) * (kt**2 * Psi0 / 15) * dl
        if self.w0wa:
            ic["dX"] = -sp.Rational(3, 2) * (1 + S["wX"]) * Psi0
            ic["tX"] = sp.Rational(1, 2) * k**2 * tau * Psi0
        Pi0 = Pi.subs({U[n]: ic[n] for n in names})
        ic["Phi"] = Psi0 + 12 * PI * a**2 * Pi0 / k**2
        self.ic = [ic[n] for n in names]

        # ---------------- matter overdensity Δm = Σρ_sΔ_s/Σρ_s over c, b, h (fourier.jl:39-52)
        Iu = sum(self.Ws[i] * self.xs[i] * psi[i][1] for i in range(nx))
        Idr = sum(self.Ws[i] * E[i] * psi[i][0] for i in range(nx))
        delta_h = Idr / S["Irho"]
        theta_h = k * Iu / (S["Irho"] + S["IP"] / 3)
        w_h = S["Ph"] / S["rh"]
        Dc = U["dc"] + 3 * Hc * U["tc"] / k**2
        Db = U["db"] + 3 * Hc * U["tb"] / k**2
        Dh = delta_h + 3 * Hc * (1 + w_h) * theta_h / k**2
        self.Delta_m = (S["rc"] * Dc + S["rb"] * Db + S["rh"] * Dh) / (S["rc"] + S["rb"] + S["rh"])
