// =====================================================================================
// TEST INFRASTRUCTURE ONLY -- CPU ORACLE.  Nothing in the product path may include,
// link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it (through oracle/sbref.py, ctypes).
//
// A hand-written CPU restatement of the SymBoltz.jl hot path (reference = hersle/SymBoltz.jl
// v1.6.0, file:line citations relative to /root/reference):
//   background + RECFAST thermodynamics ODE      src/models/{gravity,baryons,photons,neutrinos,
//                                                generic_species,dark_energy,cosmologies}.jl
//   "today" callback, bg Hermite spline          src/solve.jl:158-202, src/utils.jl:118-127
//   perturbation RHS (Newtonian gauge)           same model files; flat form docs/src/LCDM.md:82-325
//   initial conditions                           cosmologies.jl:74-77, generic_species.jl:54-57,
//                                                photons.jl:47-67, neutrinos.jl:33-39,134-139
//   Rodas5P + PI controller + dense output       OrdinaryDiffEqRosenbrock (NOT in the reference tree,
//                                                Project.toml:26; restated from the published method,
//                                                G. Steinebach, BIT 63 (2023), and the package's
//                                                documented controller defaults -- see DESIGN.md)
//   source functions ST, SE, matter Δm           cosmologies.jl:99-104, fourier.jl:39-52
//
// It is deliberately independent of the product's sympy code generator: equations are typed by
// hand as C++ templates and differentiated by complex-step (time gradient) and by truncated
// Taylor "jets" (source-function time derivatives); the Jacobian is probed (the system is linear in u; columns with
// disjoint row patterns are probed together).  Linear algebra is Gaussian elimination with threshold partial pivoting and zero
// skipping in a fill-reducing order (stand-in for KLU, src/solve.jl:329).
//
// PARITY PINNING: Julia is not available, so the reference itself cannot be run.  This oracle is
// pinned against the reference's own golden vectors test/class_Pk.dat and test/class_Cl.dat
// (test/runtests.jl:872-886) at the reference's own tolerances -- see tests/test_oracle_golden.py.
// Step-level behaviour of OrdinaryDiffEq is "parity unpinned" (no fixture exists).
// The alternative integrators TRBDF2 and KenCarp4 (reference ptalg(accuracy = 0 / 1), src/solve.jl:333-337) are restatements of the PUBLISHED schemes
// (OrdinaryDiffEqSDIRK is an un-vendored dependency): "parity unpinned" against Julia's step selection; they are anchored on the CLASS-pinned Rodas5P
// result (convergence at the methods' orders, tests/test_host_cpu.py) and on the tableau's order conditions.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "jet.hpp"

using cplx = std::complex<double>;
static inline double realpart(double x) { return x; }
static inline double realpart(const cplx& x) { return x.real(); }
static inline double realpart(const Jet& x) { return x.c[0]; }

// ------------------------------------------------------------------ constants
// CODATA 2018 / IAU values, as used by reference src/constants.jl:3-24
namespace K {
const double pi = 3.14159265358979323846;
const double c = 299792458.0;
const double h = 6.62607015e-34;
const double hbar = h / (2 * pi);
const double kB = 1.380649e-23;
const double GN = 6.67430e-11;
const double sigmaSB = 5.670374419e-8;
const double aR = 4 / c * sigmaSB;
const double sigmaT = 6.6524587321e-29;
const double Mpc = 3.0856775814913673e22;
const double km = 1e3;
const double H100 = 100 * km / Mpc;
const double k0 = H100 * Mpc / c; // h/Mpc in H0/c units = 1/2997.92458
const double eV = 1.602176634e-19;
const double me = 9.1093837015e-31;
const double u_amu = 1.66053906660e-27;
const double mH = 1.008 * u_amu;      // PeriodicTable.jl elements[:H].atomic_mass
const double mHe = 4.0026022 * u_amu; // PeriodicTable.jl elements[:He].atomic_mass
// transitions, src/constants.jl:26-45
const double lamH_inf_1s = 91.17534e-9, EH_inf_1s = h * c / lamH_inf_1s;
const double lamH_2s_1s = 121.56700e-9, EH_2s_1s = h * c / lamH_2s_1s;
const double EH_inf_2s = EH_inf_1s - EH_2s_1s;
const double lamHe_inf_1s = 50.42590e-9, EHe_inf_1s = h * c / lamHe_inf_1s;
const double lamHe_2s_1s = 60.14045e-9, EHe_2s_1s = h * c / lamHe_2s_1s;
const double lamHe_2p_1s = 58.43344e-9, fHe_2p_1s = c / lamHe_2p_1s, EHe_2p_1s = h * fHe_2p_1s;
const double EHe_2p_2s = EHe_2p_1s - EHe_2s_1s;
const double EHe_inf_2s = EHe_inf_1s - EHe_2s_1s;
const double EHep_inf_1s = 54.4178 * eV;
const double lamHet_inf_2s = 260.0463e-9, EHet_inf_2s = h * c / lamHet_inf_2s;
const double lamHet_2p_1s = 59.1411e-9, fHet_2p_1s = c / lamHet_2p_1s, EHet_2p_1s = h * fHet_2p_1s;
const double lamHet_2s_1s = 62.5563e-9, EHet_2s_1s = h * c / lamHet_2s_1s;
const double EHet_2p_2s = EHet_2p_1s - EHet_2s_1s;
const double LambdaH = 8.2245809, LambdaHe = 51.3;    // baryons.jl:40-41
const double A2ps = 1.798287e9, A2pt = 177.58e0;      // baryons.jl:90-91
} // namespace K

#define NXMAX 32
extern "C" {
struct SboParams { // mirrors reference parameters (src/parameters.jl:3-19) + model switches
    double h, Omega_c, Omega_b, T0, YHe, Neff, m_eV, Nh, ln_As1e10, ns, w0, wa, cs2X;
    int lmax, nx, model; // model: 0 = ΛCDM, 1 = w0waCDM (src/models/cosmologies.jl:211-215)
    int pad;
    double x[NXMAX], W[NXMAX]; // momentum nodes/weights, neutrinos.jl:55-60 (computed in sbref.py)
};
}

struct Derived {
    SboParams p;
    double H0SI, Omega_g, Omega_nu, Omega_h, Omega_L, fHe, y0, Irho0, Ch, Tnu0;
    double dlnf0[NXMAX];
    double kappa0 = 0, tau0 = 0;
    int N; // number of perturbation unknowns
    // state layout
    int iPhi, idc, itc, idb, itb, iF, iG, iN, ipsi, idX, itX;
};

static void derive(const SboParams& p, Derived& D) {
    using namespace K;
    D.p = p;
    D.H0SI = H100 * p.h;
    D.Omega_g = pi * pi / 15 * std::pow(kB * p.T0, 4) / (std::pow(hbar, 3) * std::pow(c, 5)) * 8 * pi * GN / (3 * D.H0SI * D.H0SI); // photons.jl:44-46
    D.Tnu0 = std::cbrt(4.0 / 11.0) * p.T0;                                                                                           // cosmologies.jl:79
    D.Omega_nu = p.Neff * 7.0 / 8.0 * std::pow(4.0 / 11.0, 4.0 / 3.0) * D.Omega_g;                                                   // cosmologies.jl:80
    double m = p.m_eV * eV / (c * c);
    D.y0 = m * c * c / (kB * D.Tnu0); // neutrinos.jl:81
    D.Irho0 = 0;
    for (int i = 0; i < p.nx; i++) {
        D.Irho0 += p.W[i] * std::sqrt(p.x[i] * p.x[i] + D.y0 * D.y0);
        D.dlnf0[i] = -p.x[i] / (1 + std::exp(-p.x[i])); // neutrinos.jl:73
    }
    D.Omega_h = p.Nh * 8 * pi / 3 * 2 / (2 * pi * pi) * std::pow(kB * D.Tnu0, 4) / std::pow(hbar * c, 3) * D.Irho0 / (std::pow(D.H0SI * c, 2) / GN); // neutrinos.jl:83
    D.Omega_L = 1 - (D.Omega_g + D.Omega_nu + p.Omega_c + p.Omega_b + D.Omega_h);                                                                       // cosmologies.jl:221-227
    D.fHe = p.YHe / (mHe / mH * (1 - p.YHe));                                                                                                           // baryons.jl:151
    D.Ch = 3 / (8 * pi) * D.Omega_h / D.Irho0;
    int L = p.lmax, nh = L + 1;
    D.iPhi = 0; D.idc = 1; D.itc = 2; D.idb = 3; D.itb = 4;
    D.iF = 5; D.iG = D.iF + nh; D.iN = D.iG + nh; D.ipsi = D.iN + nh;
    D.idX = D.ipsi + p.nx * nh; D.itX = D.idX + 1;
    D.N = D.idX + (p.model == 1 ? 2 : 0);
}

// smoothifelse, src/utils.jl:37
template <class T> static inline T smoothifelse(const T& x, const T& v1, const T& v2, double k) {
    using std::tanh;
    return 0.5 * ((v1 + v2) + (v2 - v1) * tanh(k * x));
}

// ------------------------------------------------------------------ background chain
template <class T> struct Bg {
    T a, adot, Hc, rho_c, rho_b, rho_g, rho_nu, rho_h, P_h, rho_X, w_X, wdot_X, rho, P;
    T Tg, Tb, DTg, DTb, nH, Xe, ne, kappadot, muc2, csb2;
    T E[NXMAX], Irho, IP;
    T dXH, dXHe, dDT; // bg RHS (only if full)
};

// y = (a, _κ, XH⁺, XHe⁺, ΔT)
template <class T> static void bg_chain(const Derived& D, const T* y, Bg<T>& o, bool full) {
    using namespace K;
    using std::exp; using std::pow; using std::sqrt; using std::tanh; using std::log;
    const SboParams& p = D.p;
    const T a = y[0], XH = y[2], XHe = y[3], DT = y[4];
    const double pre = 3 / (8 * pi);
    T a2 = a * a, a3 = a2 * a, a4 = a2 * a2;
    o.a = a;
    o.rho_c = pre * p.Omega_c / a3; // generic_species.jl:35-36
    o.rho_b = pre * p.Omega_b / a3;
    o.rho_g = pre * D.Omega_g / a4;
    o.rho_nu = pre * D.Omega_nu / a4;
    // massive neutrinos, neutrinos.jl:107-115,126
    T y_h = D.y0 * a;
    o.Irho = T(0.0); o.IP = T(0.0);
    for (int i = 0; i < p.nx; i++) {
        o.E[i] = sqrt(p.x[i] * p.x[i] + y_h * y_h);
        o.Irho += p.W[i] * o.E[i];
        o.IP += (p.W[i] * p.x[i] * p.x[i]) / o.E[i];
    }
    o.rho_h = D.Ch * o.Irho / a4;
    o.P_h = D.Ch * o.IP / (3.0 * a4);
    if (p.model == 1) { // dark_energy.jl:45-50 (analytical w0wa)
        o.w_X = p.w0 + p.wa * (1.0 - a);
        o.rho_X = pre * D.Omega_L * pow(a, -3 * (1 + p.w0 + p.wa)) * exp(-3 * p.wa * (1.0 - a));
    } else { // cosmological constant, dark_energy.jl:6-17
        o.w_X = T(-1.0);
        o.rho_X = T(pre * D.Omega_L);
    }
    o.rho = o.rho_c + o.rho_b + o.rho_g + o.rho_nu + o.rho_h + o.rho_X;            // cosmologies.jl:86
    o.P = o.rho_g / 3.0 + o.rho_nu / 3.0 + o.P_h + o.w_X * o.rho_X;                // cosmologies.jl:87
    o.adot = sqrt(8 * pi / 3 * o.rho) * a2;                                        // gravity.jl:31
    o.Hc = o.adot / a;                                                             // metric.jl:24
    o.wdot_X = -p.wa * o.adot;
    // thermodynamics, baryons.jl:145-200
    o.Tg = p.T0 / a;
    o.DTg = -o.Tg * o.Hc;
    o.Tb = DT + o.Tg;
    o.nH = (1 - p.YHe) * o.rho_b * (D.H0SI * D.H0SI) / GN / mH; // baryons.jl:190
    T nHe = D.fHe * o.nH;
    T beta = 1.0 / (kB * o.Tb);
    T lame = h / sqrt(2 * pi * me / beta); // baryons.jl:45
    T lame3 = lame * lame * lame;
    // He++ Saha, baryons.jl:67-68
    T RHe = exp(-beta * EHep_inf_1s) / (o.nH * lame3);
    T den = 1.0 + D.fHe + RHe;
    T XHepp = 2.0 * RHe * D.fHe / den / (1.0 + sqrt(1.0 + 4.0 * RHe * D.fHe / (den * den)));
    // reionization, baryons.jl:128-138,202-203
    T opz = 1.0 / a; // 1+z
    T Xre1, Xre2;
    {
        double z1 = 7.6711, dz = 0.5, n = 1.5;
        double f1 = std::sqrt(1 + z1) * (1 + z1);
        Xre1 = smoothifelse(f1 - sqrt(opz) * opz, T(0.0), T(1 + D.fHe), 1 / (n * std::pow(1 + z1, n - 1) * dz));
        double z2 = 3.5;
        Xre2 = smoothifelse((1 + z2) - opz, T(0.0), T(D.fHe), 1 / (1.0 * dz));
    }
    o.Xe = XH + D.fHe * XHe + XHepp + Xre1 + Xre2; // baryons.jl:70,205
    o.ne = o.Xe * o.nH;
    o.kappadot = -a / D.H0SI * o.ne * sigmaT * c;        // baryons.jl:176
    o.muc2 = mH * c * c / (1 + (mH / mHe - 1) * p.YHe + o.Xe * (1 - p.YHe)); // baryons.jl:183
    o.DTb = -2.0 * o.Tb * o.Hc - a / p.h * (8.0 / 3.0 * sigmaT * aR / H100) * (o.Tg * o.Tg * o.Tg * o.Tg) / (me * c) * o.Xe / (1.0 + D.fHe + o.Xe) * DT; // baryons.jl:185
    o.csb2 = kB / o.muc2 * (o.Tb - o.DTb / (3.0 * o.Hc)); // baryons.jl:182
    o.dDT = o.DTb - o.DTg;                                 // baryons.jl:187
    if (!full) return;
    // ---- RECFAST rate equations, baryons.jl:43-116
    T HSI = D.H0SI * (o.Hc / a); // baryons.jl:46
    T Tr = o.Tb / 1e4;
    T alphaH = 1.125 * 1e-19 * 4.309 * pow(Tr, -0.6166) / (1.0 + 0.6703 * pow(Tr, 0.5300)); // baryons.jl:43,49,76
    T betaH = alphaH / lame3 * exp(-beta * EH_inf_2s);
    T lna = log(a);
    T g1 = (lna + 7.28) / 0.18, g2 = (lna + 6.73) / 0.33;
    T KHfit = 1.0 + (-0.14) * exp(-(g1 * g1)) + 0.079 * exp(-(g2 * g2)); // baryons.jl:77-78
    T KH = KHfit / (8 * pi) * (lamH_2s_1s * lamH_2s_1s * lamH_2s_1s) / HSI;
    T CHfull = (1.0 + KH * LambdaH * o.nH * (1.0 - XH)) / (1.0 + KH * (LambdaH + betaH) * o.nH * (1.0 - XH));
    T CH = smoothifelse(XH - 0.99, CHfull, T(1.0), 1e3);
    o.dXH = -a / D.H0SI * CH * (alphaH * XH * o.ne - betaH * (1.0 - XH) * exp(-beta * EH_2s_1s)); // baryons.jl:53
    // He singlet
    auto alphaHefit = [&](double q, double pp) {
        const double T1 = std::pow(10.0, 5.114), T2 = 3.0;
        T s2 = sqrt(o.Tb / T2), s1 = sqrt(o.Tb / T1);
        return q / (s2 * pow(1.0 + s2, 1 - pp) * pow(1.0 + s1, 1 + pp)); // baryons.jl:44
    };
    const double eps = 1e-9;
    T alphaHe = alphaHefit(std::pow(10.0, -16.744), 0.711);
    T betaHe = 4.0 * alphaHe / lame3 * exp(-beta * EHe_inf_2s);
    T invKHe0 = 8 * pi * HSI / (lamHe_2p_1s * lamHe_2p_1s * lamHe_2p_1s);
    T tauHe = 3 * A2ps * nHe * (1.0 - XHe + eps) / invKHe0;
    T invKHe1 = -exp(-tauHe) * invKHe0;
    auto gammaHe = [&](double A, double sig, double f) {
        return 3 * A * D.fHe * (1.0 - XHe + eps) * (c * c) / (8 * pi * sig * sqrt(2 * pi / (beta * mHe * c * c)) * (1.0 - XH + eps) * (f * f * f)); // baryons.jl:92
    };
    T g2ps = gammaHe(A2ps, 1.436289e-22, fHe_2p_1s);
    T invKHe2 = A2ps / (1.0 + 0.36 * pow(g2ps, 0.86)) * 3.0 * nHe * (1.0 - XHe);
    T KHe = 1.0 / (invKHe0 + invKHe1 + invKHe2);
    T e2p2s = exp(-beta * EHe_2p_2s);
    T CHefull = (e2p2s + KHe * LambdaHe * nHe * (1.0 - XHe)) / (e2p2s + KHe * (LambdaHe + betaHe) * nHe * (1.0 - XHe));
    T CHe = smoothifelse(XHe - 0.99, CHefull, T(1.0), 1e3);
    T DXHes = -a / D.H0SI * CHe * (alphaHe * XHe * o.ne - betaHe * (1.0 - XHe) * exp(-beta * EHe_2s_1s));
    // He triplet, baryons.jl:103-110
    T alphaHet = alphaHefit(std::pow(10.0, -16.306), 0.761);
    T betaHet = 4.0 / 3.0 * alphaHet / lame3 * exp(-beta * EHet_inf_2s);
    T tauHet = A2pt * nHe * (1.0 - XHe + eps) * 3.0 * (lamHet_2p_1s * lamHet_2p_1s * lamHet_2p_1s) / (8 * pi * HSI);
    T pHet = (1.0 - exp(-tauHet)) / tauHet;
    T g2pt = gammaHe(A2pt, 1.484872e-22, fHet_2p_1s);
    T CHetnum = A2pt * (pHet + 1.0 / (1.0 + 0.66 * pow(g2pt, 0.9)) / 3.0) * exp(-beta * EHet_2p_2s);
    T CHet = (eps + CHetnum) / (eps + CHetnum + betaHet);
    T DXHet = -a / D.H0SI * CHet * (alphaHet * XHe * o.ne - betaHet * (1.0 - XHe) * 3.0 * exp(-beta * EHet_2s_1s));
    o.dXHe = DXHes + DXHet; // baryons.jl:64
}

template <class T> static void bg_rhs(const Derived& D, const T* y, T* dy) {
    Bg<T> o;
    bg_chain(D, y, o, true);
    dy[0] = o.adot; dy[1] = o.kappadot; dy[2] = o.dXH; dy[3] = o.dXHe; dy[4] = o.dDT;
}

// ------------------------------------------------------------------ Rodas5P tableau
namespace R5 {
const double gamma = 0.21193756319429014;
const double a21 = 3.0;
const double a31 = 2.849394379747939, a32 = 0.45842242204463923;
const double a41 = -6.954028509809101, a42 = 2.489845061869568, a43 = -10.358996098473584;
const double a51 = 2.8029986275628964, a52 = 0.5072464736228206, a53 = -0.3988312541770524, a54 = -0.04721187230404641;
const double a61 = -7.502846399306121, a62 = 2.561846144803919, a63 = -11.627539656261098, a64 = -0.18268767659942256, a65 = 0.030198172008377946;
const double C21 = -14.155112264123755;
const double C31 = -17.97296035885952, C32 = -2.859693295451294;
const double C41 = 147.12150275711716, C42 = -1.41221402718213, C43 = 71.68940251302358;
const double C51 = 165.43517024871676, C52 = -0.4592823456491126, C53 = 42.90938336958603, C54 = -5.961986721573306;
const double C61 = 24.854864614690072, C62 = -3.0009227002832186, C63 = 47.4931110020768, C64 = 5.5814197821558125, C65 = -0.6610691825249471;
const double C71 = 30.91273214028599, C72 = -3.1208243349937974, C73 = 77.79954646070892, C74 = 34.28646028294783, C75 = -19.097331116725623, C76 = -28.087943162872662;
const double C81 = 37.80277123390563, C82 = -3.2571969029072276, C83 = 112.26918849496327, C84 = 66.9347231244047, C85 = -40.06618937091002, C86 = -54.66780262877968, C87 = -9.48861652309627;
const double c2 = 0.6358126895828704, c3 = 0.4095798393397535, c4 = 0.9769306725060716, c5 = 0.4288403609558664;
const double d1 = 0.21193756319429014, d2 = -0.42387512638858027, d3 = -0.3384627126235924, d4 = 1.8046452872882734, d5 = 2.325825639765069;
const double H2[8] = {25.948786856663858, -2.5579724845846235, 10.433815404888879, -2.3679251022685204, 0.524948541321073, 1.1241088310450404, 0.4272876194431874, -0.17202221070155493};
const double H3[8] = {-9.91568850695171, -0.9689944594115154, 3.0438037242978453, -24.495224566215796, 20.176138334709044, 15.98066361424651, -6.789040303419874, -6.710236069923372};
const double H4[8] = {11.419903575922262, 2.8879645146136994, 72.92137995996029, 80.12511834622643, -52.072871366152654, -59.78993625266729, -0.15582684282751913, 4.883087185713722};
const double A[8][8] = {
    {0}, {a21}, {a31, a32}, {a41, a42, a43}, {a51, a52, a53, a54}, {a61, a62, a63, a64, a65},
    {a61, a62, a63, a64, a65, 1.0}, {a61, a62, a63, a64, a65, 1.0, 1.0}};
const double C[8][8] = {
    {0}, {C21}, {C31, C32}, {C41, C42, C43}, {C51, C52, C53, C54}, {C61, C62, C63, C64, C65},
    {C71, C72, C73, C74, C75, C76}, {C81, C82, C83, C84, C85, C86, C87}};
const double cc[8] = {0, c2, c3, c4, c5, 1, 1, 1};
const double dd[8] = {d1, d2, d3, d4, d5, 0, 0, 0};
} // namespace R5

// dense LU with partial pivoting (row-major n×n)
static bool lu_factor(int n, double* A, int* piv) {
    for (int k = 0; k < n; k++) {
        int p = k; double m = std::fabs(A[k * n + k]);
        for (int i = k + 1; i < n; i++) { double v = std::fabs(A[i * n + k]); if (v > m) { m = v; p = i; } }
        piv[k] = p;
        if (!(m > 0)) return false;
        if (p != k) for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[p * n + j]);
        double inv = 1 / A[k * n + k];
        for (int i = k + 1; i < n; i++) {
            double l = A[i * n + k] * inv;
            if (l != 0) { A[i * n + k] = l; double* Ai = A + i * n; const double* Ak = A + k * n; for (int j = k + 1; j < n; j++) Ai[j] -= l * Ak[j]; }
        }
    }
    return true;
}
static void lu_solve(int n, const double* A, const int* piv, double* b) {
    for (int k = 0; k < n; k++) { if (piv[k] != k) std::swap(b[k], b[piv[k]]); }
    for (int i = 1; i < n; i++) { double s = b[i]; const double* Ai = A + i * n; for (int j = 0; j < i; j++) s -= Ai[j] * b[j]; b[i] = s; }
    for (int i = n - 1; i >= 0; i--) { double s = b[i]; const double* Ai = A + i * n; for (int j = i + 1; j < n; j++) s -= Ai[j] * b[j]; b[i] = s / Ai[i]; }
}

// Sparse-aware LU on a dense array: threshold partial pivoting (keep the diagonal unless it is < 1e-3 of the column maximum --
// KLU's default pivot tolerance, the linear solver the reference selects at src/solve.jl:329), zero skipping in the elimination
// and index lists for the triangular solves.  With the fill-reducing ordering used by PtSys (hierarchy tails first) the cost is
// O(nnz(L+U)) instead of O(n^3); the arithmetic is ordinary Gaussian elimination.
struct SparseLU {
    int n = 0;
    std::vector<double> A;
    std::vector<int> piv, Lptr, Lidx, Uptr, Uidx, cols;
    bool factor(int n_, const double* W) {
        n = n_; A.assign(W, W + (size_t)n * n); piv.resize(n); cols.resize(n);
        for (int k = 0; k < n; k++) {
            int p = k; double m = std::fabs(A[(size_t)k * n + k]), mx = m;
            for (int i = k + 1; i < n; i++) { double v = std::fabs(A[(size_t)i * n + k]); if (v > mx) { mx = v; p = i; } }
            if (!(mx > 0)) return false;
            if (m >= 1e-3 * mx) p = k; // diagonal preference
            piv[k] = p;
            if (p != k) for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
            int nc = 0; const double* Ak = &A[(size_t)k * n];
            for (int j = k + 1; j < n; j++) if (Ak[j] != 0) cols[nc++] = j;
            const double inv = 1 / Ak[k];
            for (int i = k + 1; i < n; i++) {
                double* Ai = &A[(size_t)i * n];
                if (Ai[k] != 0) { const double l = Ai[k] * inv; Ai[k] = l; for (int c = 0; c < nc; c++) Ai[cols[c]] -= l * Ak[cols[c]]; }
            }
        }
        Lptr.assign(n + 1, 0); Uptr.assign(n + 1, 0); Lidx.clear(); Uidx.clear();
        for (int i = 0; i < n; i++) {
            const double* Ai = &A[(size_t)i * n];
            for (int j = 0; j < i; j++) if (Ai[j] != 0) Lidx.push_back(j);
            Lptr[i + 1] = (int)Lidx.size();
            for (int j = i + 1; j < n; j++) if (Ai[j] != 0) Uidx.push_back(j);
            Uptr[i + 1] = (int)Uidx.size();
        }
        return true;
    }
    void solve(double* b) const {
        for (int k = 0; k < n; k++) if (piv[k] != k) std::swap(b[k], b[piv[k]]);
        for (int i = 0; i < n; i++) { double s = b[i]; const double* Ai = &A[(size_t)i * n]; for (int q = Lptr[i]; q < Lptr[i + 1]; q++) s -= Ai[Lidx[q]] * b[Lidx[q]]; b[i] = s; }
        for (int i = n - 1; i >= 0; i--) { double s = b[i]; const double* Ai = &A[(size_t)i * n]; for (int q = Uptr[i]; q < Uptr[i + 1]; q++) s -= Ai[Uidx[q]] * b[Uidx[q]]; b[i] = s / Ai[i]; }
    }
};

enum { RC_SUCCESS = 0, RC_MAXITERS = 1, RC_DTMIN = 2, RC_UNSTABLE = 3, RC_TERMINATED = 4 };
struct Stats { long naccept = 0, nreject = 0, nf = 0, njac = 0; };

// One Rosenbrock step workspace + adaptive driver.  Sys provides n, f(t,u,du), jac(t,u,J), tgrad(t,u,dT).
template <class Sys> struct Rodas5P {
    Sys& sys; int n;
    std::vector<double> W, Wp, ks, U, du, dT, rhs, f0, K1, K2, K3, unew, tmp;
    std::vector<int> perm; // elimination order (fill-reducing), identity if the system provides none
    SparseLU lu;
    Stats st;
    explicit Rodas5P(Sys& s) : sys(s), n(s.n), W(n * n), Wp(n * n), ks(8 * n), U(n), du(n), dT(n), rhs(n), f0(n), K1(n), K2(n), K3(n), unew(n), tmp(n), perm(n) {
        for (int i = 0; i < n; i++) perm[i] = i;
        sys.ordering(perm.data());
    }
    void solve_inplace(double* b) { // W x = b with the symmetrically permuted factorisation
        for (int i = 0; i < n; i++) tmp[i] = b[perm[i]];
        lu.solve(tmp.data());
        for (int i = 0; i < n; i++) b[perm[i]] = tmp[i];
    }

    // one step of size dt from (t,u); fills unew, ks, K1..K3. returns false if LU singular.
    bool step(double t, const double* u, double dt) {
        using namespace R5;
        sys.f(t, u, f0.data()); st.nf++;
        sys.tgrad(t, u, dT.data());
        sys.jac(t, u, W.data()); st.njac++;
        double dg = 1 / (R5::gamma * dt);
        for (int i = 0; i < n; i++) { const double* Wi = &W[(size_t)perm[i] * n]; double* Pi = &Wp[(size_t)i * n]; for (int j = 0; j < n; j++) Pi[j] = -Wi[perm[j]]; Pi[i] += dg; }
        if (!lu.factor(n, Wp.data())) return false;
        for (int s = 0; s < 8; s++) {
            double* k = ks.data() + s * n;
            const double* fs;
            if (s == 0) { fs = f0.data(); }
            else {
                if (s <= 5) { for (int i = 0; i < n; i++) { double v = u[i]; for (int j = 0; j < s; j++) v += A[s][j] * ks[j * n + i]; U[i] = v; } }
                else { for (int i = 0; i < n; i++) U[i] += ks[(s - 1) * n + i]; }
                sys.f(t + cc[s] * dt, U.data(), du.data()); st.nf++;
                fs = du.data();
            }
            for (int i = 0; i < n; i++) { double v = fs[i] + dt * dd[s] * dT[i]; for (int j = 0; j < s; j++) v += (C[s][j] / dt) * ks[j * n + i]; k[i] = v; }
            solve_inplace(k);
        }
        for (int i = 0; i < n; i++) unew[i] = U[i] + ks[7 * n + i];
        for (int i = 0; i < n; i++) {
            double s1 = 0, s2 = 0, s3 = 0;
            for (int j = 0; j < 8; j++) { double kj = ks[j * n + i]; s1 += H2[j] * kj; s2 += H3[j] * kj; s3 += H4[j] * kj; }
            K1[i] = s1; K2[i] = s2; K3[i] = s3;
        }
        return true;
    }
    double errnorm(const double* u, double abstol, double reltol) const {
        double s = 0; const double* k8 = ks.data() + 7 * n;
        for (int i = 0; i < n; i++) { double sc = abstol + reltol * std::max(std::fabs(u[i]), std::fabs(unew[i])); double r = k8[i] / sc; s += r * r; }
        return std::sqrt(s / n);
    }
    // dense output at theta in [0,1] of the last computed step
    void interp(const double* u0, const double* u1, double th, double* out) const {
        double t1 = 1 - th;
        for (int i = 0; i < n; i++) out[i] = t1 * u0[i] + th * (u1[i] + t1 * (K1[i] + th * (K2[i] + th * K3[i])));
    }
    void interp_deriv(const double* u0, const double* u1, double th, double dt, double* out) const {
        for (int i = 0; i < n; i++) out[i] = (K1[i] + th * (-2 * K1[i] + 2 * K2[i] + th * (-3 * K2[i] + 3 * K3[i] - 4 * th * K3[i])) - u0[i] + u1[i]) / dt;
    }
    // Hairer-style automatic initial step (OrdinaryDiffEq initdt; assumption recorded in DESIGN.md)
    double initdt(double t0, const double* u0, double tend, double abstol, double reltol) {
        std::vector<double> sk(n), fa(n), fb(n), u1(n);
        sys.f(t0, u0, fa.data()); st.nf++;
        double d0 = 0, d1 = 0;
        for (int i = 0; i < n; i++) { sk[i] = abstol + std::fabs(u0[i]) * reltol; d0 += (u0[i] / sk[i]) * (u0[i] / sk[i]); d1 += (fa[i] / sk[i]) * (fa[i] / sk[i]); }
        d0 = std::sqrt(d0 / n); d1 = std::sqrt(d1 / n);
        double dtmax = tend - t0, smalldt = 1e-6;
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? smalldt : 0.01 * d0 / d1;
        dt0 = std::min(dt0, dtmax);
        for (int i = 0; i < n; i++) u1[i] = u0[i] + dt0 * fa[i];
        sys.f(t0 + dt0, u1.data(), fb.data()); st.nf++;
        double d2 = 0;
        for (int i = 0; i < n; i++) { double r = (fb[i] - fa[i]) / sk[i]; d2 += r * r; }
        d2 = std::sqrt(d2 / n) / dt0;
        double dm = std::max(d1, d2);
        double dt1 = (dm <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2 + std::log10(dm)) / 5.0);
        return std::min(std::min(100 * dt0, dt1), dtmax);
    }
};

struct Controller { // PI controller, OrdinaryDiffEq defaults for an order-5 adaptive method
    double beta1 = 7.0 / 50, beta2 = 2.0 / 25, gamma = 0.9, qmin = 0.2, qmax = 10, qold = 1e-4, q11 = 1;
    double q_of(double EEst) {
        if (EEst == 0) return 1 / qmax;
        q11 = std::pow(EEst, beta1);
        double q = q11 / std::pow(qold, beta2);
        return std::max(1 / qmax, std::min(1 / qmin, q / gamma));
    }
    double accept(double dt, double q, double EEst) { qold = std::max(EEst, 1e-4); return dt / q; }
    double reject(double dt) { return dt / std::min(1 / qmin, q11 / gamma); }
};

// TRBDF2 (Bank et al. 1985; Hosea & Shampine 1996): the reference's `ptalg(prob; accuracy = 0)` (src/solve.jl:333-335) hands the perturbations to
// OrdinaryDiffEq.jl's TRBDF2 with NLNewton(κ = 1).  OrdinaryDiffEq is an un-vendored dependency (absent from /root/reference) and there is no Julia
// here, so this restates the PUBLISHED scheme -- PARITY UNPINNED against Julia's step selection.  The system is linear in u, so each implicit stage is
// solved exactly as one linear system with J at the stage time (a converged Newton iteration), not with a frozen, reused W.
//   γ = 2 − √2, d = γ/2, ω = √2/4;  stage 2 (trapezoidal rule to t + γ dt):  (I − d dt J_γ) u_γ = u_n + d dt f_n
//   stage 3 (BDF2 to t + dt):  (I − d dt J_1) u_{n+1} = u_γ/(γ(2−γ)) − (1−γ)² u_n/(γ(2−γ));   f of a stage follows from its own equation (FSAL)
//   error estimate: dt Σ (b_i − b̂_i) f_i with the third-order companion b̂ = ((1−ω)/3, (3ω+1)/3, d/3), filtered through (I − d dt J_1)^{-1}
//   dense output: cubic Hermite on (u_n, f_n, u_{n+1}, f_{n+1}).
template <class Sys> struct TRBDF2 {
    Sys& sys; int n;
    std::vector<double> W, Wp, f0, fg, f1, ug, unew, est, tmp, rhs;
    std::vector<int> perm;
    SparseLU lu;
    Stats st;
    bool have_f0 = false;
    explicit TRBDF2(Sys& s) : sys(s), n(s.n), W(n * n), Wp(n * n), f0(n), fg(n), f1(n), ug(n), unew(n), est(n), tmp(n), rhs(n), perm(n) {
        for (int i = 0; i < n; i++) perm[i] = i;
        sys.ordering(perm.data());
    }
    void solve_inplace(double* b) {
        for (int i = 0; i < n; i++) tmp[i] = b[perm[i]];
        lu.solve(tmp.data());
        for (int i = 0; i < n; i++) b[perm[i]] = tmp[i];
    }
    bool factor(double ts, const double* u, double x) { // x I − J(ts), symmetrically permuted
        sys.jac(ts, u, W.data()); st.njac++;
        for (int i = 0; i < n; i++) { const double* Wi = &W[(size_t)perm[i] * n]; double* Pi = &Wp[(size_t)i * n]; for (int j = 0; j < n; j++) Pi[j] = -Wi[perm[j]]; Pi[i] += x; }
        return lu.factor(n, Wp.data());
    }
    bool step(double t, const double* u, double dt) {
        const double g = 2 - std::sqrt(2.0), d = g / 2, w = std::sqrt(2.0) / 4, x = 1 / (d * dt);
        const double a = 1 / (g * (2 - g)), b = (1 - g) * (1 - g) / (g * (2 - g));
        if (!have_f0) { sys.f(t, u, f0.data()); st.nf++; have_f0 = true; }
        if (!factor(t + g * dt, u, x)) return false;
        for (int i = 0; i < n; i++) ug[i] = x * u[i] + f0[i];
        solve_inplace(ug.data());
        for (int i = 0; i < n; i++) fg[i] = x * (ug[i] - u[i]) - f0[i];
        if (!factor(t + dt, u, x)) return false;
        for (int i = 0; i < n; i++) { rhs[i] = a * ug[i] - b * u[i]; unew[i] = x * rhs[i]; }
        solve_inplace(unew.data());
        for (int i = 0; i < n; i++) f1[i] = x * (unew[i] - rhs[i]);
        const double e1 = (4 * w - 1) / 3, e2 = -1.0 / 3, e3 = 2 * d / 3;
        for (int i = 0; i < n; i++) est[i] = x * (dt * (e1 * f0[i] + e2 * fg[i] + e3 * f1[i]));
        solve_inplace(est.data());
        return true;
    }
    void accept() { f0 = f1; } // first-same-as-last
    double errnorm(const double* u, double abstol, double reltol) const {
        double s = 0;
        for (int i = 0; i < n; i++) { double sc = abstol + reltol * std::max(std::fabs(u[i]), std::fabs(unew[i])); double r = est[i] / sc; s += r * r; }
        return std::sqrt(s / n);
    }
    void interp(const double* u0, double th, double dt, double* out) const {
        for (int i = 0; i < n; i++) {
            const double du = unew[i] - u0[i];
            out[i] = (1 - th) * u0[i] + th * unew[i] + th * (th - 1) * ((1 - 2 * th) * du + (th - 1) * dt * f0[i] + th * dt * f1[i]);
        }
    }
};
// KenCarp4 (Kennedy & Carpenter 2003, ARK4(3)6L[2]SA; the reference's `ptalg(prob; accuracy = 1)`, src/solve.jl:336-337, and the algorithm its sparse-Jacobian
// test solves with, test/runtests.jl:580-590).  The reference integrates the whole right-hand side implicitly, so only the ESDIRK half of the pair matters:
// γ = 1/4, six stages, stiffly accurate, third-order companion b̂.  Same caveats as TRBDF2: published tableau (row sums Σ_j a_ij = c_i and Σ b̂ = 1 are
// asserted in tests/test_host_cpu.py), stages solved exactly with J at the stage time, PARITY UNPINNED against OrdinaryDiffEq.jl's step selection.
namespace KC4 {
const double g = 0.25;
const double c[6] = {0, 0.5, 83.0 / 250, 31.0 / 50, 17.0 / 20, 1};
const double A[6][5] = {{0, 0, 0, 0, 0},
                        {0.25, 0, 0, 0, 0},
                        {8611.0 / 62500, -1743.0 / 31250, 0, 0, 0},
                        {5012029.0 / 34652500, -654441.0 / 2922500, 174375.0 / 388108, 0, 0},
                        {15267082809.0 / 155376265600.0, -71443401.0 / 120774400, 730878875.0 / 902184768, 2285395.0 / 8070912, 0},
                        {82889.0 / 524892, 0, 15625.0 / 83664, 69875.0 / 102672, -2260.0 / 8211}};
const double b[6] = {82889.0 / 524892, 0, 15625.0 / 83664, 69875.0 / 102672, -2260.0 / 8211, 0.25};
const double bh[6] = {4586570599.0 / 29645900160.0, 0, 178811875.0 / 945068544, 814220225.0 / 1159782912, -3700637.0 / 11593932, 61727.0 / 225920};
}
template <class Sys> struct KenCarp4 {
    Sys& sys; int n;
    std::vector<double> W, Wp, F, f0, f1, U, unew, est, tmp, rhs;
    std::vector<int> perm;
    SparseLU lu;
    Stats st;
    bool have_f0 = false;
    explicit KenCarp4(Sys& s) : sys(s), n(s.n), W(n * n), Wp(n * n), F(6 * n), f0(n), f1(n), U(n), unew(n), est(n), tmp(n), rhs(n), perm(n) {
        for (int i = 0; i < n; i++) perm[i] = i;
        sys.ordering(perm.data());
    }
    void solve_inplace(double* b) {
        for (int i = 0; i < n; i++) tmp[i] = b[perm[i]];
        lu.solve(tmp.data());
        for (int i = 0; i < n; i++) b[perm[i]] = tmp[i];
    }
    bool factor(double ts, const double* u, double x) {
        sys.jac(ts, u, W.data()); st.njac++;
        for (int i = 0; i < n; i++) { const double* Wi = &W[(size_t)perm[i] * n]; double* Pi = &Wp[(size_t)i * n]; for (int j = 0; j < n; j++) Pi[j] = -Wi[perm[j]]; Pi[i] += x; }
        return lu.factor(n, Wp.data());
    }
    bool step(double t, const double* u, double dt) {
        const double x = 1 / (KC4::g * dt);
        if (!have_f0) { sys.f(t, u, f0.data()); st.nf++; have_f0 = true; }
        std::copy(f0.begin(), f0.end(), F.begin());
        for (int s = 1; s < 6; s++) {
            for (int i = 0; i < n; i++) { double v = 0; for (int j = 0; j < s; j++) v += KC4::A[s][j] * F[(size_t)j * n + i]; rhs[i] = u[i] + dt * v; U[i] = x * rhs[i]; }
            if (!factor(t + KC4::c[s] * dt, u, x)) return false;
            solve_inplace(U.data());
            for (int i = 0; i < n; i++) F[(size_t)s * n + i] = x * (U[i] - rhs[i]);
        }
        unew = U; // stiffly accurate
        for (int i = 0; i < n; i++) { f1[i] = F[(size_t)5 * n + i]; double v = 0; for (int j = 0; j < 6; j++) v += (KC4::b[j] - KC4::bh[j]) * F[(size_t)j * n + i]; est[i] = x * (dt * v); }
        solve_inplace(est.data());
        return true;
    }
    void accept() { f0 = f1; }
    double errnorm(const double* u, double abstol, double reltol) const {
        double s = 0;
        for (int i = 0; i < n; i++) { double sc = abstol + reltol * std::max(std::fabs(u[i]), std::fabs(unew[i])); double r = est[i] / sc; s += r * r; }
        return std::sqrt(s / n);
    }
    void interp(const double* u0, double th, double dt, double* out) const {
        for (int i = 0; i < n; i++) {
            const double du = unew[i] - u0[i];
            out[i] = (1 - th) * u0[i] + th * unew[i] + th * (th - 1) * ((1 - 2 * th) * du + (th - 1) * dt * f0[i] + th * dt * f1[i]);
        }
    }
};
// Step-size control of the implicit methods: Gustafsson's predictive controller in the form OrdinaryDiffEq uses for its Newton-based methods
// (restated from the published controller, same caveat as above); exponent 1/3 for TRBDF2's O(dt³) estimate, one "Newton iteration" per stage.
struct PredictiveController {
    double gamma = 0.9, qmin = 0.2, qmax = 10, qsteady_min = 1, qsteady_max = 1.2, expo = 1.0 / 3;
    double qold = 1, dtacc = 0, erracc = 0; long success_iter = 0;
    double q_of(double EEst) {
        if (EEst == 0) return qold = 1 / qmax;
        return qold = std::max(1 / qmax, std::min(1 / qmin, std::pow(EEst, expo) / gamma));
    }
    double accept(double dt, double q, double EEst) {
        double qacc = q;
        if (success_iter > 0) {
            double qgus = (dtacc / dt) * std::pow(EEst * EEst / erracc, expo);
            qgus = std::max(1 / qmax, std::min(1 / qmin, qgus / gamma));
            qacc = std::max(q, qgus);
        }
        if (qsteady_min <= qacc && qacc <= qsteady_max) qacc = 1;
        success_iter++;
        dtacc = dt; erracc = std::max(1e-2, EEst);
        return dt / qacc;
    }
    double reject(double dt) { return success_iter == 0 ? 0.1 * dt : dt / qold; }
};

// ------------------------------------------------------------------ background system and solve
struct BgSys {
    const Derived& D; int n = 5;
    explicit BgSys(const Derived& d) : D(d) {}
    void ordering(int*) {}
    void f(double, const double* u, double* du) { bg_rhs<double>(D, u, du); }
    void tgrad(double, const double*, double* dT) { for (int i = 0; i < 5; i++) dT[i] = 0; } // autonomous
    void jac(double, const double* u, double* J) { // complex-step columns (exact to rounding)
        const double e = 1e-30;
        for (int j = 0; j < 5; j++) {
            cplx y[5], dy[5];
            for (int i = 0; i < 5; i++) y[i] = u[i];
            y[j] += cplx(0, e);
            bg_rhs<cplx>(D, y, dy);
            for (int i = 0; i < 5; i++) J[i * 5 + j] = dy[i].imag() / e;
        }
    }
};

struct BgSolution {
    std::vector<double> t, y, dy; // knots, values (nb×5), derivatives (nb×5)
    double tau0 = 0, kappa0 = 0, taurec = 0; int retcode = RC_SUCCESS; Stats st;
};

static double solve_a_ini(const Derived& D, double tau) { // ℋ = 1/τ initially, cosmologies.jl:76
    double a = std::sqrt(D.Omega_g + D.Omega_nu) * tau;
    for (int it = 0; it < 100; it++) {
        auto F = [&](double aa) { double y[5] = {aa, 0, 1, 1, 0}; Bg<double> o; bg_chain(D, y, o, false); return o.adot * tau / aa - 1.0; };
        double f = F(a), h = a * 1e-7, fp = (F(a + h) - F(a - h)) / (2 * h);
        double an = a - f / fp;
        if (std::fabs(an - a) <= 1e-15 * std::fabs(a)) { a = an; break; }
        a = an;
    }
    return a;
}

static void solve_background(Derived& D, double tini, double tmax, double reltol, double abstol, BgSolution& S) {
    BgSys sys(D);
    Rodas5P<BgSys> R(sys);
    Controller ctl;
    double u[5] = {solve_a_ini(D, tini), 0, 1, 1, 0}; // baryons.jl:28,33,156,164
    double t = tini;
    std::vector<std::vector<double>> K1s, K2s, K3s; std::vector<double> dts;
    S.t.clear(); S.y.clear(); S.dy.clear();
    S.t.push_back(t); S.y.insert(S.y.end(), u, u + 5);
    double dt = R.initdt(t, u, tmax, abstol, reltol);
    double d_first[5]; bool have_first = false;
    for (long it = 0;; it++) {
        if (it >= 100000) { S.retcode = RC_MAXITERS; break; }
        if (t + dt > tmax) dt = tmax - t;
        if (!R.step(t, u, dt)) { S.retcode = RC_UNSTABLE; break; }
        double EEst = R.errnorm(u, abstol, reltol);
        if (!std::isfinite(EEst)) { // NaN stage (e.g. pow of a tiny negative number in RECFAST): reject and shrink by 1/qmin
            R.st.nreject++; dt /= 5; if (dt < 1e-14 * t) { S.retcode = RC_UNSTABLE; break; } continue; }
        double q = ctl.q_of(EEst);
        if (EEst > 1) { R.st.nreject++; dt = ctl.reject(dt); if (dt < 1e-14 * t) { S.retcode = RC_DTMIN; break; } continue; }
        R.st.naccept++;
        double dtnew = ctl.accept(dt, q, EEst);
        const double* un = R.unew.data();
        if (!have_first) { R.interp_deriv(u, un, 0.0, dt, d_first); have_first = true; S.dy.insert(S.dy.end(), d_first, d_first + 5); }
        if (un[0] - 1.0 >= 0) {
            // "today" callback: a crosses 1 inside this step (solve.jl:158-202). Right-root bisection on the dense output.
            double lo = 0, hi = 1, tmp[5];
            for (int b = 0; b < 200; b++) {
                double mid = 0.5 * (lo + hi);
                if (mid == lo || mid == hi) break;
                R.interp(u, un, mid, tmp);
                if (tmp[0] - 1.0 >= 0) hi = mid; else lo = mid;
            }
            double uend[5]; R.interp(u, un, hi, uend);
            double dtr = hi * dt, tr = t + dtr;
            // re-evaluate the dense-output vectors for the shortened step (OrdinaryDiffEq addsteps! after change_t_via_interpolation!)
            double dend[5];
            if (R.step(t, u, dtr)) R.interp_deriv(u, uend, 1.0, dtr, dend); else bg_rhs<double>(D, uend, dend);
            S.t.push_back(tr); S.y.insert(S.y.end(), uend, uend + 5); S.dy.insert(S.dy.end(), dend, dend + 5);
            S.tau0 = tr; S.kappa0 = uend[1]; S.retcode = RC_SUCCESS;
            break;
        }
        double dn[5]; R.interp_deriv(u, un, 1.0, dt, dn);
        t = t + dt; for (int i = 0; i < 5; i++) u[i] = un[i];
        S.t.push_back(t); S.y.insert(S.y.end(), u, u + 5); S.dy.insert(S.dy.end(), dn, dn + 5);
        if (t >= tmax) { S.tau0 = t; S.kappa0 = u[1]; break; }
        dt = dtnew;
    }
    S.st = R.st;
    D.tau0 = S.tau0; D.kappa0 = S.kappa0;
    // τrec = knot with maximal visibility v = -κ̇ exp(-κ) over solver steps (solve.jl:183-189)
    double vmax = -1; S.taurec = S.tau0;
    for (size_t i = 0; i < S.t.size(); i++) { Bg<double> o; bg_chain(D, &S.y[5 * i], o, false); double v = -o.kappadot * std::exp(-(S.y[5 * i + 1] - S.kappa0)); if (v > vmax) { vmax = v; S.taurec = S.t[i]; } }
}

// Cubic Hermite spline of the 5 bg unknowns (utils.jl:118-127; DataInterpolations CubicHermiteSpline)
struct Spline {
    int nb; const double *t, *y, *dy;
    int find(double tr) const { // interval index idx with t[idx] <= tr <= t[idx+1], clamped (Extension extrapolation)
        int lo = 0, hi = nb - 1;
        while (hi - lo > 1) { int mid = (lo + hi) / 2; if (t[mid] <= tr) lo = mid; else hi = mid; }
        return lo;
    }
    template <class T> void eval(const T& tau, T* out) const {
        int i = find(realpart(tau));
        double h = t[i + 1] - t[i];
        T d0 = tau - t[i], d1 = tau - t[i + 1];
        for (int j = 0; j < 5; j++) {
            double u0 = y[5 * i + j], u1 = y[5 * i + 5 + j], v0 = dy[5 * i + j], v1 = dy[5 * i + 5 + j];
            double c1 = (u1 - u0 - v0 * h) / (h * h), c2 = (v1 - v0 - 2 * c1 * h) / (h * h);
            out[j] = u0 + d0 * v0 + d0 * d0 * (c1 + d1 * c2);
        }
    }
};

// ------------------------------------------------------------------ perturbations
// Observed perturbation quantities given bg coefficients and state.
template <class T> struct PtObs { T drho, Pi, Psi, Phidot, thetag, Pig; };

template <class T> static void pt_rhs(const Derived& D, const Bg<T>& c, const T& tau, double k, const T* u, T* du, PtObs<T>* obs = nullptr) {
    using namespace K;
    const SboParams& p = D.p;
    const int L = p.lmax, nh = L + 1, nx = p.nx;
    const T *F = u + D.iF, *G = u + D.iG, *Nn = u + D.iN;
    T *dF = du + D.iF, *dG = du + D.iG, *dN = du + D.iN;
    const T Phi = u[D.iPhi], dc = u[D.idc], tc = u[D.itc], db = u[D.idb], tb = u[D.itb];
    T a2 = c.a * c.a, a4 = a2 * a2;
    // massive neutrino moments, neutrinos.jl:117-124
    T Idrho = T(0.0), Isig = T(0.0);
    for (int i = 0; i < nx; i++) {
        const T* psi = u + D.ipsi + i * nh;
        Idrho += p.W[i] * c.E[i] * psi[0];
        Isig += (p.W[i] * p.x[i] * p.x[i]) / c.E[i] * psi[2];
    }
    T drho_h = D.Ch / a4 * Idrho;                   // δh ρh
    T Pi_h = D.Ch / a4 * (2.0 / 3.0) * Isig;        // (ρh+Ph) σh
    T drho = dc * c.rho_c + db * c.rho_b + F[0] * c.rho_g + Nn[0] * c.rho_nu + drho_h; // cosmologies.jl:92
    if (p.model == 1) drho += u[D.idX] * c.rho_X;
    T Pi = (4.0 / 3.0) * c.rho_g * (F[2] / 2.0) + (4.0 / 3.0) * c.rho_nu * (Nn[2] / 2.0) + Pi_h; // cosmologies.jl:94
    T Psi = Phi - 12 * pi * a2 * Pi / (k * k);                                                  // gravity.jl:39
    T Phidot = -4 * pi / 3 * a2 / c.Hc * drho - (k * k) / (3.0 * c.Hc) * Phi - c.Hc * Psi;       // gravity.jl:38
    du[D.iPhi] = Phidot;
    // cold dark matter, generic_species.jl:46-47 (w = cs² = 0)
    du[D.idc] = -(tc - 3.0 * Phidot);
    du[D.itc] = -c.Hc * tc + (k * k) * Psi;
    // baryons (w = 0, thermal cs², Thomson drag), generic_species.jl:46-47, cosmologies.jl:95
    T thetag = 3.0 * k * F[1] / 4.0;
    T tint = -c.kappadot * 4.0 * c.rho_g / (3.0 * c.rho_b) * (thetag - tb);
    du[D.idb] = -(tb - 3.0 * Phidot) - 3.0 * c.Hc * c.csb2 * db;
    du[D.itb] = -c.Hc * tb + c.csb2 * (k * k) * db + (k * k) * Psi + tint;
    // photons, photons.jl:28-60
    T Pig = F[2] + G[0] + G[2];
    dF[0] = -k * F[1] + 4.0 * Phidot;
    dF[1] = k / 3 * (F[0] - 2.0 * F[2] + 4.0 * Psi) - 4.0 / 3.0 * c.kappadot / k * (tb - thetag);
    for (int l = 2; l < L; l++) dF[l] = k / (2 * l + 1) * (double(l) * F[l - 1] - double(l + 1) * F[l + 1]) + c.kappadot * (F[l] - (l == 2 ? 1.0 : 0.0) * Pig / 10.0);
    dF[L] = k * F[L - 1] - double(L + 1) / tau * F[L] + c.kappadot * F[L];
    dG[0] = k * (-G[1]) + c.kappadot * (G[0] - Pig / 2.0);
    dG[1] = k / 3 * (G[0] - 2.0 * G[2]) + c.kappadot * G[1];
    for (int l = 2; l < L; l++) dG[l] = k / (2 * l + 1) * (double(l) * G[l - 1] - double(l + 1) * G[l + 1]) + c.kappadot * (G[l] - (l == 2 ? 1.0 : 0.0) * Pig / 10.0);
    dG[L] = k * G[L - 1] - double(L + 1) / tau * G[L] + c.kappadot * G[L];
    // massless neutrinos, neutrinos.jl:22-26
    dN[0] = -k * Nn[1] + 4.0 * Phidot;
    dN[1] = k / 3 * (Nn[0] - 2.0 * Nn[2] + 4.0 * Psi);
    for (int l = 2; l < L; l++) dN[l] = k / (2 * l + 1) * (double(l) * Nn[l - 1] - double(l + 1) * Nn[l + 1]);
    dN[L] = k * Nn[L - 1] - double(L + 1) / tau * Nn[L];
    // massive neutrinos, neutrinos.jl:128-132
    for (int i = 0; i < nx; i++) {
        const T* psi = u + D.ipsi + i * nh; T* dpsi = du + D.ipsi + i * nh;
        T xE = p.x[i] / c.E[i], Ex = c.E[i] / p.x[i];
        dpsi[0] = -k * xE * psi[1] - Phidot * D.dlnf0[i];
        dpsi[1] = k / 3 * xE * (psi[0] - 2.0 * psi[2]) - k / 3 * Ex * Psi * D.dlnf0[i];
        for (int l = 2; l < L; l++) dpsi[l] = k / (2 * l + 1) * xE * (double(l) * psi[l - 1] - double(l + 1) * psi[l + 1]);
        dpsi[L] = k / (2 * L + 1) * xE * (double(L) * psi[L - 1] - double(L + 1) * (double(2 * L + 1) * Ex * psi[L] / (k * tau) - psi[L - 1]));
    }
    if (p.model == 1) { // w0wa fluid, dark_energy.jl:53-59
        T w = c.w_X, dX = u[D.idX], tX = u[D.itX];
        T ca2 = w - c.wdot_X / (3.0 * c.Hc * (1.0 + w));
        double cs2 = p.cs2X;
        du[D.idX] = -(1.0 + w) * (tX - 3.0 * Phidot) - 3.0 * c.Hc * (cs2 - w) * dX - 9.0 * (c.Hc / k) * (c.Hc / k) * (1.0 + w) * (cs2 - ca2) * tX;
        du[D.itX] = -c.Hc * (1 - 3 * cs2) * tX + cs2 / (1.0 + w) * (k * k) * dX + (k * k) * Psi;
    }
    if (obs) { obs->drho = drho; obs->Pi = Pi; obs->Psi = Psi; obs->Phidot = Phidot; obs->thetag = thetag; obs->Pig = Pig; }
}

// closed-form adiabatic initial conditions (see file header for reference lines)
static void pt_initial(const Derived& D, const Spline& spl, double tau, double k, double* u) {
    using namespace K;
    const SboParams& p = D.p; const int L = p.lmax, nh = L + 1;
    double y[5]; spl.eval(tau, y);
    Bg<double> c; bg_chain(D, y, c, false);
    for (int i = 0; i < D.N; i++) u[i] = 0;
    double fnu = (c.rho_nu + c.rho_h) / (c.rho_g + c.rho_nu + c.rho_h); // cosmologies.jl:89
    double Psi = 20 * 0.5 / (15 + 4 * fnu);                             // cosmologies.jl:55,75
    double kt = k * tau, kd = k / c.kappadot;
    u[D.idc] = -1.5 * Psi; u[D.itc] = 0.5 * k * k * tau * Psi;
    u[D.idb] = -1.5 * Psi; u[D.itb] = 0.5 * k * k * tau * Psi;
    double *F = u + D.iF, *G = u + D.iG, *Nn = u + D.iN;
    F[0] = -2 * Psi; F[1] = 2.0 / 3.0 * kt * Psi; F[2] = -8.0 / 15.0 * kd * F[1]; F[3] = -3.0 / 7.0 * kd * F[2];
    G[0] = 5.0 / 16.0 * F[2]; G[1] = -1.0 / 16.0 * kd * F[2]; G[2] = 1.0 / 16.0 * F[2]; G[3] = -3.0 / 7.0 * kd * G[2];
    double thnu = 0.5 * k * k * tau * Psi, signu = kt * kt * Psi / 15;
    Nn[0] = -2 * Psi; Nn[1] = 4 * thnu / (3 * k); Nn[2] = 2 * signu; Nn[3] = 3.0 / 7.0 * kt * Nn[2];
    double Isig = 0;
    for (int i = 0; i < p.nx; i++) {
        double* psi = u + D.ipsi + i * nh; double dl = D.dlnf0[i];
        psi[0] = -0.25 * (-2 * Psi) * dl;
        psi[1] = -1.0 / 3.0 * c.E[i] / p.x[i] * (0.5 * kt * Psi) * dl;
        psi[2] = -0.5 * (kt * kt * Psi / 15) * dl;
        Isig += p.W[i] * p.x[i] * p.x[i] / c.E[i] * psi[2];
    }
    if (p.model == 1) { u[D.idX] = -1.5 * (1 + c.w_X) * Psi; u[D.itX] = 0.5 * k * k * tau * Psi; } // dark_energy.jl:61-64
    double a2 = c.a * c.a, a4 = a2 * a2;
    double Pi = (4.0 / 3.0) * c.rho_g * F[2] / 2 + (4.0 / 3.0) * c.rho_nu * Nn[2] / 2 + D.Ch / a4 * (2.0 / 3.0) * Isig;
    u[D.iPhi] = Psi + 12 * pi * a2 * Pi / (k * k); // gravity.jl:39
}

struct PtSys {
    const Derived& D; const Spline& spl; double k; int n;
    std::vector<double> e, col;
    std::vector<int> color, rowsptr, rowsidx; // column colouring for compressed Jacobian probing + per-column row pattern
    int ncolors = 0;
    PtSys(const Derived& d, const Spline& s, double kk) : D(d), spl(s), k(kk), n(d.N), e(d.N), col(d.N) {}
    // fill-reducing elimination order: hierarchy tails (l = lmax ... 3 of every hierarchy) first, the coupled core last
    void ordering(int* perm) {
        const int L = D.p.lmax, nh = L + 1; int q = 0;
        std::vector<char> used(n, 0);
        auto tail = [&](int base) { for (int l = L; l >= 3; l--) { perm[q++] = base + l; used[base + l] = 1; } };
        tail(D.iF); tail(D.iG); tail(D.iN);
        for (int i = 0; i < D.p.nx; i++) tail(D.ipsi + i * nh);
        for (int i = 0; i < n; i++) if (!used[i]) perm[q++] = i;
    }
    void f(double t, const double* u, double* du) {
        double y[5]; spl.eval(t, y);
        Bg<double> c; bg_chain(D, y, c, false);
        pt_rhs<double>(D, c, t, k, u, du);
    }
    // The system is linear in u: column j of J is f(e_j).  The sparsity pattern is fixed, so it is discovered once by full
    // probing (at two times) and afterwards columns with disjoint row patterns are probed together (greedy colouring).
    void discover(double t) {
        std::vector<std::vector<int>> rows(n);
        double ts[2] = {t, 0.5 * (t + D.tau0)};
        for (int it = 0; it < 2; it++) {
            double y[5]; spl.eval(ts[it], y);
            Bg<double> c; bg_chain(D, y, c, false);
            for (int j = 0; j < n; j++) {
                std::fill(e.begin(), e.end(), 0.0); e[j] = 1;
                pt_rhs<double>(D, c, ts[it], k, e.data(), col.data());
                for (int i = 0; i < n; i++) if (col[i] != 0 && std::find(rows[j].begin(), rows[j].end(), i) == rows[j].end()) rows[j].push_back(i);
            }
        }
        color.assign(n, -1); ncolors = 0;
        std::vector<std::vector<char>> occ; // occ[color][row]
        for (int j = 0; j < n; j++) {
            int cidx = -1;
            for (int cc2 = 0; cc2 < ncolors && cidx < 0; cc2++) { bool ok = true; for (int i : rows[j]) if (occ[cc2][i]) { ok = false; break; } if (ok) cidx = cc2; }
            if (cidx < 0) { cidx = ncolors++; occ.emplace_back(n, 0); }
            color[j] = cidx; for (int i : rows[j]) occ[cidx][i] = 1;
        }
        rowsptr.assign(n + 1, 0); rowsidx.clear();
        for (int j = 0; j < n; j++) { for (int i : rows[j]) rowsidx.push_back(i); rowsptr[j + 1] = (int)rowsidx.size(); }
    }
    void jac(double t, const double*, double* J) {
        if (ncolors == 0) discover(t);
        double y[5]; spl.eval(t, y);
        Bg<double> c; bg_chain(D, y, c, false);
        std::fill(J, J + (size_t)n * n, 0.0);
        for (int cc2 = 0; cc2 < ncolors; cc2++) {
            for (int j = 0; j < n; j++) e[j] = (color[j] == cc2) ? 1.0 : 0.0;
            pt_rhs<double>(D, c, t, k, e.data(), col.data());
            for (int j = 0; j < n; j++) if (color[j] == cc2) for (int q = rowsptr[j]; q < rowsptr[j + 1]; q++) J[(size_t)rowsidx[q] * n + j] = col[rowsidx[q]];
        }
    }
    void tgrad(double t, const double* u, double* dT) { // ∂f/∂τ at fixed u through the spline (complex step == ForwardDiff dual)
        const double eps = 1e-30;
        cplx tc(t, eps), y[5];
        spl.eval(tc, y);
        Bg<cplx> c; bg_chain(D, y, c, false);
        std::vector<cplx> uc(n), duc(n);
        for (int i = 0; i < n; i++) uc[i] = u[i];
        pt_rhs<cplx>(D, c, tc, k, uc.data(), duc.data());
        for (int i = 0; i < n; i++) dT[i] = duc[i].imag() / eps;
    }
};

// Solve one mode. saveat may be empty (then only the final state is returned in uend).
// the same driver for the ESDIRK methods (TRBDF2, KenCarp4): initial step as for Rodas5P with the method's order
template <class Stepper> static int solve_mode_sdirk(const Derived& D, const Spline& spl, double k, double tini, double tend, double reltol, double abstol,
                                                     long maxiters, int nsave, const double* saveat, double* usave, double* uend, long* stats, int order);
static int solve_mode(const Derived& D, const Spline& spl, double k, double tini, double tend, double reltol, double abstol,
                      long maxiters, int nsave, const double* saveat, double* usave /*nsave×N*/, double* uend /*N*/, long* stats /*4*/, int alg = 0) {
    if (alg == 1) return solve_mode_sdirk<TRBDF2<PtSys>>(D, spl, k, tini, tend, reltol, abstol, maxiters, nsave, saveat, usave, uend, stats, 2);
    if (alg == 2) return solve_mode_sdirk<KenCarp4<PtSys>>(D, spl, k, tini, tend, reltol, abstol, maxiters, nsave, saveat, usave, uend, stats, 4);
    PtSys sys(D, spl, k);
    int n = sys.n;
    Rodas5P<PtSys> R(sys);
    Controller ctl;
    std::vector<double> u(n);
    int rc = RC_SUCCESS, isave = 0;
    if (!(k > 0) || !std::isfinite(k)) { rc = RC_UNSTABLE; for (int i = 0; i < n; i++) uend[i] = NAN; for (long i = 0; i < (long)nsave * n; i++) usave[i] = NAN; if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0; return rc; }
    pt_initial(D, spl, tini, k, u.data());
    double t = tini;
    while (isave < nsave && saveat[isave] <= t) { if (saveat[isave] == t) std::copy(u.begin(), u.end(), usave + (long)isave * n); else for (int i = 0; i < n; i++) usave[(long)isave * n + i] = NAN; isave++; }
    if (tend > tini) {
        double dt = R.initdt(t, u.data(), tend, abstol, reltol);
        for (long it = 0;; it++) {
            if (it >= maxiters) { rc = RC_MAXITERS; break; }
            bool last = false;
            if (t + dt >= tend - 100 * 2.2e-16 * std::fabs(tend)) { dt = tend - t; last = true; }
            if (!R.step(t, u.data(), dt)) { rc = RC_UNSTABLE; break; }
            double EEst = R.errnorm(u.data(), abstol, reltol);
            if (getenv("SBO_TRACE")) fprintf(stderr, "TRACE %ld %.17g %.17g %.17g\n", it, t, dt, EEst);
            if (!std::isfinite(EEst)) { R.st.nreject++; dt /= 5; if (dt < 1e-15 * std::fabs(t)) { rc = RC_UNSTABLE; break; } continue; }
            double q = ctl.q_of(EEst);
            if (EEst > 1) { R.st.nreject++; dt = ctl.reject(dt); if (dt < 1e-15 * std::fabs(t)) { rc = RC_DTMIN; break; } continue; }
            R.st.naccept++;
            double dtnew = ctl.accept(dt, q, EEst);
            double tn = last ? tend : t + dt;
            while (isave < nsave && saveat[isave] <= tn) {
                double* out = usave + (long)isave * n;
                if (saveat[isave] == tn) std::copy(R.unew.begin(), R.unew.end(), out);
                else R.interp(u.data(), R.unew.data(), (saveat[isave] - t) / dt, out);
                isave++;
            }
            t = tn; u = R.unew;
            bool bad = false; for (int i = 0; i < n; i++) if (std::isnan(u[i])) bad = true;
            if (bad) { rc = RC_UNSTABLE; break; }
            if (last) break;
            dt = dtnew;
        }
    }
    std::copy(u.begin(), u.end(), uend);
    for (; isave < nsave; isave++) for (int i = 0; i < n; i++) usave[(long)isave * n + i] = NAN;
    if (stats) { stats[0] = R.st.naccept; stats[1] = R.st.nreject; stats[2] = R.st.nf; stats[3] = R.st.njac; }
    return rc;
}

template <class Stepper> static int solve_mode_sdirk(const Derived& D, const Spline& spl, double k, double tini, double tend, double reltol, double abstol,
                                                     long maxiters, int nsave, const double* saveat, double* usave, double* uend, long* stats, int order) {
    PtSys sys(D, spl, k);
    int n = sys.n;
    Stepper R(sys);
    PredictiveController ctl;
    ctl.expo = 1.0 / (order + (order == 2 ? 1 : 0)); // the estimate is O(dt³) for TRBDF2 (third-order companion of a second-order method), O(dt⁴) for KenCarp4 (4(3) pair)
    std::vector<double> u(n);
    int rc = RC_SUCCESS, isave = 0;
    if (!(k > 0) || !std::isfinite(k)) { for (int i = 0; i < n; i++) uend[i] = NAN; for (long i = 0; i < (long)nsave * n; i++) usave[i] = NAN; if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0; return RC_UNSTABLE; }
    pt_initial(D, spl, tini, k, u.data());
    double t = tini;
    while (isave < nsave && saveat[isave] <= t) { if (saveat[isave] == t) std::copy(u.begin(), u.end(), usave + (long)isave * n); else for (int i = 0; i < n; i++) usave[(long)isave * n + i] = NAN; isave++; }
    if (tend > tini) {
        double dt;
        { // Hairer's initial step with the method's order: same recipe as Rodas5P::initdt, exponent 1/(order + 1)
            std::vector<double> sk(n), fa(n), fb(n), u1(n);
            sys.f(t, u.data(), fa.data()); R.st.nf++;
            double d0 = 0, d1 = 0;
            for (int i = 0; i < n; i++) { sk[i] = abstol + std::fabs(u[i]) * reltol; d0 += (u[i] / sk[i]) * (u[i] / sk[i]); d1 += (fa[i] / sk[i]) * (fa[i] / sk[i]); }
            d0 = std::sqrt(d0 / n); d1 = std::sqrt(d1 / n);
            double dtmax = tend - t;
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            dt0 = std::min(dt0, dtmax);
            for (int i = 0; i < n; i++) u1[i] = u[i] + dt0 * fa[i];
            sys.f(t + dt0, u1.data(), fb.data()); R.st.nf++;
            double d2 = 0;
            for (int i = 0; i < n; i++) { double r = (fb[i] - fa[i]) / sk[i]; d2 += r * r; }
            d2 = std::sqrt(d2 / n) / dt0;
            double dm = std::max(d1, d2);
            double dt1 = (dm <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2 + std::log10(dm)) / (order + 1.0));
            dt = std::min(std::min(100 * dt0, dt1), dtmax);
        }
        for (long it = 0;; it++) {
            if (it >= maxiters) { rc = RC_MAXITERS; break; }
            bool last = false;
            if (t + dt >= tend - 100 * 2.2e-16 * std::fabs(tend)) { dt = tend - t; last = true; }
            if (!R.step(t, u.data(), dt)) { rc = RC_UNSTABLE; break; }
            double EEst = R.errnorm(u.data(), abstol, reltol);
            if (!std::isfinite(EEst)) { R.st.nreject++; dt /= 5; if (dt < 1e-15 * std::fabs(t)) { rc = RC_UNSTABLE; break; } continue; }
            double q = ctl.q_of(EEst);
            if (EEst > 1) { R.st.nreject++; dt = ctl.reject(dt); if (dt < 1e-15 * std::fabs(t)) { rc = RC_DTMIN; break; } continue; }
            R.st.naccept++;
            double dtnew = ctl.accept(dt, q, EEst);
            double tn = last ? tend : t + dt;
            while (isave < nsave && saveat[isave] <= tn) {
                double* out = usave + (long)isave * n;
                if (saveat[isave] == tn) std::copy(R.unew.begin(), R.unew.end(), out);
                else R.interp(u.data(), (saveat[isave] - t) / dt, dt, out);
                isave++;
            }
            t = tn; u = R.unew; R.accept();
            bool bad = false; for (int i = 0; i < n; i++) if (std::isnan(u[i])) bad = true;
            if (bad) { rc = RC_UNSTABLE; break; }
            if (last) break;
            dt = dtnew;
        }
    }
    std::copy(u.begin(), u.end(), uend);
    for (; isave < nsave; isave++) for (int i = 0; i < n; i++) usave[(long)isave * n + i] = NAN;
    if (stats) { stats[0] = R.st.naccept; stats[1] = R.st.nreject; stats[2] = R.st.nf; stats[3] = R.st.njac; }
    return rc;
}

// ------------------------------------------------------------------ observed quantities
// Matter overdensity Δm (fourier.jl:39-52 with modes = :m → c+b+h; generic_species.jl:48, neutrinos.jl:119-121)
static double delta_m(const Derived& D, const Spline& spl, double tau, double k, const double* u) {
    const SboParams& p = D.p; int nh = p.lmax + 1;
    double y[5]; spl.eval(tau, y);
    Bg<double> c; bg_chain(D, y, c, false);
    double Idrho = 0, Iu = 0;
    for (int i = 0; i < p.nx; i++) { const double* psi = u + D.ipsi + i * nh; Idrho += p.W[i] * c.E[i] * psi[0]; Iu += p.W[i] * p.x[i] * psi[1]; }
    double delta_h = Idrho / c.Irho;
    double theta_h = k * Iu / (c.Irho + c.IP / 3);
    double w_h = c.P_h / c.rho_h;
    double Dc = u[D.idc] + 3 * c.Hc * u[D.itc] / (k * k);
    double Db = u[D.idb] + 3 * c.Hc * u[D.itb] / (k * k);
    double Dh = delta_h + 3 * c.Hc * (1 + w_h) * theta_h / (k * k);
    return (c.rho_c * Dc + c.rho_b * Db + c.rho_h * Dh) / (c.rho_c + c.rho_b + c.rho_h);
}

// CMB sources at one (τ,k) from the state u, via order-3 jets along the flow (cosmologies.jl:99-104).
// out = {ST, SE, ST_SW, ST_ISW, ST_Doppler, ST_polarization, Sψ};  Sψ = −(Ψ+Φ)(τ−τrec)/(τ0−τrec)/(τ0−τ) for τ ≥ τrec (cosmologies.jl:105)
static void cmb_sources(const Derived& D, const Spline& spl, double taurec, double tau, double k, const double* u0, double* out) {
    const int n = D.N;
    Jet tj(tau); tj.c[1] = 1;
    // background jets by Picard iteration on the bg RHS, starting from the splined value
    double y0[5]; spl.eval(tau, y0);
    Jet y[5], dy[5];
    for (int i = 0; i < 5; i++) y[i] = Jet(y0[i]);
    for (int ord = 0; ord < 3; ord++) { bg_rhs<Jet>(D, y, dy); for (int i = 0; i < 5; i++) y[i].c[ord + 1] = dy[i].c[ord] / (ord + 1); }
    Bg<Jet> c; bg_chain(D, y, c, false);
    std::vector<Jet> u(n), du(n);
    for (int i = 0; i < n; i++) u[i] = Jet(u0[i]);
    PtObs<Jet> ob;
    for (int ord = 0; ord < 3; ord++) { pt_rhs<Jet>(D, c, tj, k, u.data(), du.data(), &ob); for (int i = 0; i < n; i++) u[i].c[ord + 1] = du[i].c[ord] / (ord + 1); }
    pt_rhs<Jet>(D, c, tj, k, u.data(), du.data(), &ob);
    Jet kappa = y[1] - D.kappa0;           // baryons.jl:178
    Jet Eexp = exp(-kappa);                // baryons.jl:179
    Jet v = deriv(Eexp);                   // baryons.jl:180
    Jet Phi = u[D.iPhi], Psi = ob.Psi;
    Jet ub = u[D.itb] / k;
    Jet F0 = u[D.iF];
    double ST_SW = (v * (F0 / 4.0 + Psi + ob.Pig / 16.0)).c[0];
    double ST_ISW = (Eexp * (deriv(Psi) + deriv(Phi))).c[0];
    double ST_Dop = deriv(v * ub).c[0] / k;
    double ST_pol = 3 / (16 * k * k) * deriv(deriv(v * ob.Pig)).c[0];
    double chi = D.tau0 - tau;
    double SE = 3.0 / 16.0 * (v * ob.Pig).c[0] / ((k * chi) * (k * chi));
    out[0] = ST_SW + ST_ISW + ST_Dop + ST_pol; out[1] = SE; out[2] = ST_SW; out[3] = ST_ISW; out[4] = ST_Dop; out[5] = ST_pol;
    out[6] = (tau >= taurec) ? -(Psi.c[0] + Phi.c[0]) * (tau - taurec) / (D.tau0 - taurec) / (D.tau0 - tau) : 0.0;
}

// ------------------------------------------------------------------ C interface (ctypes)
extern "C" {

int sbo_nstate(const SboParams* p) { Derived D; derive(*p, D); return D.N; }

// derived parameters: out = {Omega_g, Omega_nu, Omega_h, Omega_L, fHe, y0, Irho0, kpivot, As}
void sbo_derived(const SboParams* p, double* out) {
    Derived D; derive(*p, D);
    out[0] = D.Omega_g; out[1] = D.Omega_nu; out[2] = D.Omega_h; out[3] = D.Omega_L; out[4] = D.fHe; out[5] = D.y0; out[6] = D.Irho0;
    out[7] = 0.05 / K::Mpc / (K::H100 / K::c) / p->h; // inflation.jl:6
    out[8] = std::exp(p->ln_As1e10) / 1e10;           // inflation.jl:4
}

// Background solve. Returns number of knots (<= cap) or -1 if cap too small. info = {tau0, kappa0, taurec, retcode, naccept, nreject}
int sbo_background(const SboParams* p, double tini, double tmax, double reltol, double abstol, int cap, double* t, double* y, double* dy, double* info) {
    Derived D; derive(*p, D);
    BgSolution S; solve_background(D, tini, tmax, reltol, abstol, S);
    int nb = (int)S.t.size();
    info[0] = S.tau0; info[1] = S.kappa0; info[2] = S.taurec; info[3] = S.retcode; info[4] = S.st.naccept; info[5] = S.st.nreject;
    if (nb > cap) return -1;
    std::copy(S.t.begin(), S.t.end(), t); std::copy(S.y.begin(), S.y.end(), y); std::copy(S.dy.begin(), S.dy.end(), dy);
    return nb;
}

// Background observables at times taus from the spline. out[nt × 16]:
// a, ℋ, Xe, κ̇, κ, v, csb2, Tb, ρ_tot, ρ_h, P_h, fν, ȧ(spline-derivative-free), _κ, XH, XHe
void sbo_bg_observe(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double kappa0, int nt, const double* taus, double* out) {
    Derived D; derive(*p, D); D.kappa0 = kappa0;
    Spline spl{nb, t, y, dy};
    for (int i = 0; i < nt; i++) {
        double yy[5]; spl.eval(taus[i], yy);
        Bg<double> c; bg_chain(D, yy, c, true);
        double kap = yy[1] - kappa0; double* o = out + 16 * i;
        o[0] = c.a; o[1] = c.Hc; o[2] = c.Xe; o[3] = c.kappadot; o[4] = kap; o[5] = -c.kappadot * std::exp(-kap); o[6] = c.csb2; o[7] = c.Tb;
        o[8] = c.rho; o[9] = c.rho_h; o[10] = c.P_h; o[11] = (c.rho_nu + c.rho_h) / (c.rho_g + c.rho_nu + c.rho_h); o[12] = c.adot; o[13] = yy[1]; o[14] = yy[2]; o[15] = yy[3];
    }
}

void sbo_bg_rhs(const SboParams* p, const double* y, double* dy, double* J) {
    Derived D; derive(*p, D); BgSys sys(D); sys.f(0, y, dy); sys.jac(0, y, J);
}

void sbo_pt_initial(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double tau, double k, double* u) {
    Derived D; derive(*p, D); Spline spl{nb, t, y, dy}; pt_initial(D, spl, tau, k, u);
}
// f, dense J (row-major), dT at (tau, u) for one k -- used to cross-check the product's generated code
void sbo_pt_fjt(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double tau, double k, const double* u, double* f, double* J, double* dT) {
    Derived D; derive(*p, D); Spline spl{nb, t, y, dy}; PtSys sys(D, spl, k);
    sys.f(tau, u, f); sys.jac(tau, u, J); sys.tgrad(tau, u, dT);
}

// Perturbation solve over nk modes (OpenMP over modes; mirrors Threads.@spawn per mode, solve.jl:566).
// tini[nk] must already be clamped to the bg span (solve.jl:527). usave: [nk][nsave][N] or NULL, uend: [nk][N].
void sbo_solvept(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double tau0, double kappa0,
                 int nk, const double* ks, const double* tini, int nsave, const double* saveat, double reltol, double abstol, long maxiters,
                 int nthreads, double* usave, double* uend, int* retcode, long* stats) {
    Derived D; derive(*p, D); D.tau0 = tau0; D.kappa0 = kappa0;
    Spline spl{nb, t, y, dy};
    int N = D.N;
    std::vector<int> order(nk);
    for (int i = 0; i < nk; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return ks[a] > ks[b]; }); // most expensive first
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < nk; j++) {
        int i = order[j];
        std::vector<double> dummy((size_t)std::max(1, nsave) * N);
        double* us = usave ? usave + (size_t)i * nsave * N : dummy.data();
        retcode[i] = solve_mode(D, spl, ks[i], tini[i], tau0, reltol, abstol, maxiters, nsave, saveat, us, uend + (size_t)i * N, stats ? stats + 4 * i : nullptr);
    }
}

// sbo_solvept with a choice of integrator: alg 0 = Rodas5P (the reference's default, ptalg accuracy = 2), 1 = TRBDF2 (accuracy = 0), 2 = KenCarp4 (accuracy = 1)
void sbo_kencarp4_tableau(double* out /*[6][5] A, [6] c, [6] b, [6] bh, gamma*/) {
    int q = 0;
    for (int i = 0; i < 6; i++) for (int j = 0; j < 5; j++) out[q++] = KC4::A[i][j];
    for (int i = 0; i < 6; i++) out[q++] = KC4::c[i];
    for (int i = 0; i < 6; i++) out[q++] = KC4::b[i];
    for (int i = 0; i < 6; i++) out[q++] = KC4::bh[i];
    out[q++] = KC4::g;
}
void sbo_solvept_alg(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double tau0, double kappa0,
                     int nk, const double* ks, const double* tini, int nsave, const double* saveat, double reltol, double abstol, long maxiters,
                     int nthreads, double* usave, double* uend, int* retcode, long* stats, int alg) {
    Derived D; derive(*p, D); D.tau0 = tau0; D.kappa0 = kappa0;
    Spline spl{nb, t, y, dy};
    int N = D.N;
    std::vector<int> order(nk);
    for (int i = 0; i < nk; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return ks[a] > ks[b]; });
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < nk; j++) {
        int i = order[j];
        std::vector<double> dummy((size_t)std::max(1, nsave) * N);
        double* us = usave ? usave + (size_t)i * nsave * N : dummy.data();
        retcode[i] = solve_mode(D, spl, ks[i], tini[i], tau0, reltol, abstol, maxiters, nsave, saveat, us, uend + (size_t)i * N, stats ? stats + 4 * i : nullptr, alg);
    }
}

// Δm(τ,k) for states u[nk][N] at a common time tau
void sbo_delta_m(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double tau, int nk, const double* ks, const double* u, double* out) {
    Derived D; derive(*p, D); Spline spl{nb, t, y, dy};
    for (int i = 0; i < nk; i++) out[i] = delta_m(D, spl, tau, ks[i], u + (size_t)i * D.N);
}

// CMB sources on a grid: u[nk][nt][N] → out[nk][nt][7]
void sbo_sources(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double tau0, double kappa0, double taurec,
                 int nk, const double* ks, int nt, const double* taus, const double* u, double* out) {
    Derived D; derive(*p, D); D.tau0 = tau0; D.kappa0 = kappa0;
    Spline spl{nb, t, y, dy};
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < nk; i++)
        for (int j = 0; j < nt; j++) cmb_sources(D, spl, taurec, taus[j], ks[i], u + ((size_t)i * nt + j) * D.N, out + ((size_t)i * nt + j) * 7);
}

// Timing of the per-step building blocks of the CPU path (diagnostic for the reported CPU baseline): out[0..5] = µs per call of
// f, jac, tgrad, LU factor, solve, and the number of colours used by the compressed Jacobian probing.
void sbo_bench_parts(const SboParams* p, int nb, const double* t, const double* y, const double* dy, double tau, double k, double h, int reps, double* out) {
    Derived D; derive(*p, D); Spline spl{nb, t, y, dy}; PtSys sys(D, spl, k);
    int n = sys.n; Rodas5P<PtSys> R(sys);
    std::vector<double> u(n, 0.1), du(n), J((size_t)n * n), b(n, 1.0);
    auto now = [] { return omp_get_wtime(); };
    sys.jac(tau, u.data(), J.data());
    double t0 = now(); for (int i = 0; i < reps; i++) sys.f(tau, u.data(), du.data()); out[0] = (now() - t0) / reps * 1e6;
    t0 = now(); for (int i = 0; i < reps; i++) sys.jac(tau, u.data(), J.data()); out[1] = (now() - t0) / reps * 1e6;
    t0 = now(); for (int i = 0; i < reps; i++) sys.tgrad(tau, u.data(), du.data()); out[2] = (now() - t0) / reps * 1e6;
    double dg = 1 / (R5::gamma * h);
    t0 = now();
    for (int r = 0; r < reps; r++) {
        for (int i = 0; i < n; i++) { const double* Wi = &J[(size_t)R.perm[i] * n]; double* Pi = &R.Wp[(size_t)i * n]; for (int j = 0; j < n; j++) Pi[j] = -Wi[R.perm[j]]; Pi[i] += dg; }
        R.lu.factor(n, R.Wp.data());
    }
    out[3] = (now() - t0) / reps * 1e6;
    t0 = now(); for (int i = 0; i < reps; i++) { R.solve_inplace(b.data()); b[0] = 1; } out[4] = (now() - t0) / reps * 1e6;
    out[5] = sys.ncolors;
    out[6] = (double)(R.lu.Lidx.size() + R.lu.Uidx.size() + n);
}

// Generic Rodas5P self-test hooks (used by tests to re-verify the recalled tableau):
// integrates y' = f(t,y) for the 2-D nonlinear test problem with fixed steps and returns y(tend).
struct TestSys {
    int n = 2; double lam;
    void ordering(int*) {}
    void f(double t, const double* u, double* du) { du[0] = -lam * (u[0] - std::cos(t)) - std::sin(t) + 0.1 * u[1] * u[1] - 0.1 * std::sin(2 * t) * std::sin(2 * t); du[1] = 2 * std::cos(2 * t) + (u[0] - std::cos(t)); }
    void jac(double, const double* u, double* J) { J[0] = -lam; J[1] = 0.2 * u[1]; J[2] = 1; J[3] = 0; }
    void tgrad(double t, const double*, double* dT) { dT[0] = -lam * std::sin(t) - std::cos(t) - 0.4 * std::sin(2 * t) * std::cos(2 * t); dT[1] = -4 * std::sin(2 * t) + std::sin(t); }
};
// exact solution: u0 = cos t, u1 = sin 2t
void sbo_rodas_fixed(double lam, double tend, int nsteps, double* yout, double* dense_mid) {
    TestSys s; s.lam = lam; Rodas5P<TestSys> R(s);
    double u[2] = {1, 0}, t = 0, dt = tend / nsteps;
    for (int i = 0; i < nsteps; i++) {
        R.step(t, u, dt);
        if (i == nsteps - 1 && dense_mid) R.interp(u, R.unew.data(), 0.5, dense_mid);
        u[0] = R.unew[0]; u[1] = R.unew[1]; t += dt;
    }
    yout[0] = u[0]; yout[1] = u[1];
}
} // extern "C"
