// symboltz.jl_b200 -- host-side evaluation of the generated tables.  Diagnostics / unit tests of the code generator
// (f = J u, dense J, ∂f/∂τ at one point); NOT a solve path.  Compiled as plain C++ so that the generated tables are host arrays.
#include <math.h>
#include <string.h>
#include <vector>
#include "sb_model_gen.h"

struct SbSpline { int nb; const double *t, *y, *dy; };
static inline void sb_spl_eval_h(const SbSpline& s, double tau, double* y, double* yp) {
    int lo = 0, hi = s.nb - 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s.t[mid] <= tau) lo = mid; else hi = mid; }
    int i = lo;
    double h = s.t[i + 1] - s.t[i], d0 = tau - s.t[i], d1 = tau - s.t[i + 1];
    for (int j = 0; j < 5; j++) {
        double u0 = s.y[5 * i + j], u1 = s.y[5 * i + 5 + j], v0 = s.dy[5 * i + j], v1 = s.dy[5 * i + 5 + j];
        double c1 = (u1 - u0 - v0 * h) / (h * h), c2 = (v1 - v0 - 2 * c1 * h) / (h * h);
        y[j] = u0 + d0 * v0 + d0 * d0 * (c1 + d1 * c2);
        if (yp) yp[j] = v0 + 2 * d0 * (c1 + d1 * c2) + d0 * d0 * c2;
    }
}

extern "C" {
// ---- host-side evaluation of the generated tables (diagnostics / unit tests of the generator; NOT a solve path)
// f = J u, dense J (row-major N×N) and ∂f/∂τ = J' u at (τ, k) for background values y and dy/dτ = yp.
int sbm_debug_fjt(const double* P, const double* y, const double* yp, double tau, double k, const double* u, double* f, double* J, double* dT) {
    double beta[SB_NBETA], betad[SB_NBETA], b[SB_NB], bd[SB_NB];
    sb_beta(tau, y, yp, P, beta, betad);
    for (int m = 0; m < SB_NB; m++) { double kp = pow(k, (double)sb_basis_kpow[m]); b[m] = kp * beta[sb_basis_beta[m]]; bd[m] = kp * betad[sb_basis_beta[m]]; }
    for (int pass = 0; pass < 2; pass++) {
        const double* bb = pass ? bd : b; double* out = pass ? dT : f;
        std::vector<double> Jd(SB_N * SB_N, 0.0);
        for (int i = 0; i < SB_N; i++) for (int e = sb_rowptr[i]; e < sb_rowptr[i + 1]; e++) Jd[i * SB_N + sb_col[e]] += sb_coef[e] * bb[sb_bidx[e]];
        // hub outer products use b for both factors in J; for J' the product rule applies
        std::vector<double> pv(SB_N, 0), qv(SB_N, 0), ph(SB_N, 0), ps(SB_N, 0), pvd(SB_N, 0), qvd(SB_N, 0), phd(SB_N, 0), psd(SB_N, 0);
        double* vecs[4] = {pv.data(), qv.data(), ph.data(), ps.data()};
        double* vecd[4] = {pvd.data(), qvd.data(), phd.data(), psd.data()};
        for (int v = 0; v < 4; v++) for (int t = sb_hptr[v]; t < sb_hptr[v + 1]; t++) { vecs[v][sb_hidx[t]] += sb_hcoef[t] * b[sb_hb[t]]; vecd[v][sb_hidx[t]] += sb_hcoef[t] * bd[sb_hb[t]]; }
        for (int i = 0; i < SB_N; i++) for (int j = 0; j < SB_N; j++) {
            if (!pass) Jd[i * SB_N + j] += pv[i] * ph[j] + qv[i] * ps[j];
            else Jd[i * SB_N + j] += pvd[i] * ph[j] + pv[i] * phd[j] + qvd[i] * ps[j] + qv[i] * psd[j];
        }
        for (int i = 0; i < SB_N; i++) { double s = 0; for (int j = 0; j < SB_N; j++) s += Jd[i * SB_N + j] * u[j]; out[i] = s; }
        if (!pass && J) memcpy(J, Jd.data(), sizeof(double) * SB_N * SB_N);
    }
    return 0;
}
// dense J_local and the hub vectors p, q, phi, psi (each N) at (τ,k): for conditioning studies of the Woodbury split
int sbm_debug_split(const double* P, const double* y, const double* yp, double tau, double k, double* Jloc, double* hubs) {
    double beta[SB_NBETA], betad[SB_NBETA], b[SB_NB];
    sb_beta(tau, y, yp, P, beta, betad);
    for (int m = 0; m < SB_NB; m++) b[m] = pow(k, (double)sb_basis_kpow[m]) * beta[sb_basis_beta[m]];
    for (int i = 0; i < SB_N * SB_N; i++) Jloc[i] = 0;
    for (int i = 0; i < SB_N; i++) for (int e = sb_rowptr[i]; e < sb_rowptr[i + 1]; e++) Jloc[i * SB_N + sb_col[e]] += sb_coef[e] * b[sb_bidx[e]];
    for (int i = 0; i < 4 * SB_N; i++) hubs[i] = 0;
    for (int v = 0; v < 4; v++) for (int t = sb_hptr[v]; t < sb_hptr[v + 1]; t++) hubs[v * SB_N + sb_hidx[t]] += sb_hcoef[t] * b[sb_hb[t]];
    return 0;
}
int sbm_debug_initial(const double* P, const double* y, double tau, double k, double* u) { sb_initial(tau, k, y, P, u); return 0; }
int sbm_debug_bg(const double* P, const double* y, double* g, double* J, double* kder, double* obs) {
    double g2[5];
    sb_bg_rhs_jac(y, P, g, J); sb_kappa_derivs(y, P, g2, kder); sb_bg_observe(y, P, obs);
    return 0;
}
int sbm_debug_delta_m(const double* P, const double* y, double tau, double k, const double* u, double* out) { *out = sb_delta_m(tau, k, y, P, u); return 0; }
int sbm_debug_spline(int nb, const double* t, const double* y, const double* dy, double tau, double* yo, double* ypo) { SbSpline s{nb, t, y, dy}; sb_spl_eval_h(s, tau, yo, ypo); return 0; }
int sbm_debug_beta(const double* P, const double* y, const double* yp, double tau, double* beta, double* betad) { sb_beta(tau, y, yp, P, beta, betad); return 0; }
} // extern "C"
