// symboltz.jl_b200 -- per-model engine: host background solver + sm_100a perturbation kernels + C ABI.
//
// Compiled once per model (lmax, nx, w0wa) together with the generated header sb_model_gen.h
// (codegen/lower.py), mirroring the reference's problem-build-time compilation (src/solve.jl:129-236).
//
// Replaces, for the hot path only (file:line relative to the reference tree):
//   solvebg                         src/solve.jl:427-435       -> sbm_solvebg (host, Rodas5P dense, analytic J)
//   spline(bgsol)                   src/utils.jl:118-127       -> knots (t, y, y') uploaded; β-table built on device
//   setuppt/solvept, Rodas5P + KLU  src/solve.jl:496-569, 329  -> sb_integrate_kernel (one warp per k-mode, persistent,
//                                                                 atomic work queue, tables from the code generator)
//   getsym(prob.pt, Ss) sources     src/observables/fourier.jl:267-281, src/models/cosmologies.jl:99-104 -> sb_source_kernel
//   Δm for P(k)                     src/observables/fourier.jl:39-52, 90-97 -> sb_deltam_kernel
//
// There is deliberately NO CPU fallback for the perturbation path: every sbm_* device entry point returns a CUDA
// error code if no GPU is present.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "sb_model_gen.h"
#include "sb_rodas.h"

#define SB_WARP 32
#define SB_FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------ tableau
__constant__ double cA[8][8] = SB_R5_A_INIT;
__constant__ double cC[8][8] = SB_R5_C_INIT;
__constant__ double cc[8] = SB_R5_c_INIT;
__constant__ double cd[8] = SB_R5_d_INIT;
__constant__ double cH[3][8] = SB_R5_H_INIT;
__constant__ int cslot[8] = SB_R5_TIMESLOT_INIT;
static const double hA[8][8] = SB_R5_A_INIT;
static const double hC[8][8] = SB_R5_C_INIT;
static const double hc[8] = SB_R5_c_INIT;
static const double hd[8] = SB_R5_d_INIT;
static const double hH[3][8] = SB_R5_H_INIT;

enum { SB_RC_SUCCESS = 0, SB_RC_MAXITERS = 1, SB_RC_DTMIN = 2, SB_RC_UNSTABLE = 3, SB_RC_SCHEDULE = 4 };

// ------------------------------------------------------------------------------------------------ spline (host+device)
// Cubic Hermite spline of the background unknowns (reference: DataInterpolations.CubicHermiteSpline, utils.jl:126)
struct SbSpline {
    int nb;
    const double *t, *y, *dy;
};
SB_HD static inline int sb_spl_find(const SbSpline& s, double tr) {
    int lo = 0, hi = s.nb - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (s.t[mid] <= tr) lo = mid; else hi = mid;
    }
    return lo;
}
SB_HD static inline void sb_spl_eval(const SbSpline& s, double tau, double* y, double* yp) {
    int i = sb_spl_find(s, tau);
    double h = s.t[i + 1] - s.t[i], d0 = tau - s.t[i], d1 = tau - s.t[i + 1];
    for (int j = 0; j < 5; j++) {
        double u0 = s.y[5 * i + j], u1 = s.y[5 * i + 5 + j], v0 = s.dy[5 * i + j], v1 = s.dy[5 * i + 5 + j];
        double c1 = (u1 - u0 - v0 * h) / (h * h), c2 = (v1 - v0 - 2 * c1 * h) / (h * h);
        y[j] = u0 + d0 * v0 + d0 * d0 * (c1 + d1 * c2);
        if (yp) yp[j] = v0 + 2 * d0 * (c1 + d1 * c2) + d0 * d0 * c2;
    }
}

// ================================================================================================ background (host + device)
// Small dense Rodas5P for the 5 background unknowns with the analytic Jacobian from the generator
// (reference: Rodas5P + RFLUFactorization, reltol = abstol = 1e-7, src/solve.jl:313-325, 382).  The solver is one
// __host__ __device__ function: sbm_solvebg runs it on the host for one cosmology, sb_solvebg_kernel runs one thread per
// cosmology for parameter sweeps (sbm_solvebg_batch).
#ifdef __CUDA_ARCH__
#define SB_BG_A cA
#define SB_BG_C cC
#define SB_BG_H cH
#else
#define SB_BG_A hA
#define SB_BG_C hC
#define SB_BG_H hH
#endif
namespace bgsolve {
SB_HD static inline bool lu5(double* A, int* piv) {
    const int n = 5;
    for (int k = 0; k < n; k++) {
        int p = k; double m = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; i++) if (fabs(A[i * n + k]) > m) { m = fabs(A[i * n + k]); p = i; }
        piv[k] = p;
        if (!(m > 0)) return false;
        if (p != k) for (int j = 0; j < n; j++) { const double x = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = x; }
        for (int i = k + 1; i < n; i++) { double l = A[i * n + k] / A[k * n + k]; A[i * n + k] = l; for (int j = k + 1; j < n; j++) A[i * n + j] -= l * A[k * n + j]; }
    }
    return true;
}
SB_HD static inline void lusolve5(const double* A, const int* piv, double* b) {
    const int n = 5;
    for (int k = 0; k < n; k++) if (piv[k] != k) { const double x = b[k]; b[k] = b[piv[k]]; b[piv[k]] = x; }
    for (int i = 1; i < n; i++) for (int j = 0; j < i; j++) b[i] -= A[i * n + j] * b[j];
    for (int i = n - 1; i >= 0; i--) { for (int j = i + 1; j < n; j++) b[i] -= A[i * n + j] * b[j]; b[i] /= A[i * n + i]; }
}
struct Step {
    double ks[8][5], unew[5], K[3][5];
    SB_HD bool run(const double* P, const double* u, double dt) {
        double f0[5], J[25], W[25], U[5], du[5]; int piv[5];
        sb_bg_rhs_jac(u, P, f0, J);
        double dg = 1.0 / (SB_R5_GAMMA * dt);
        for (int i = 0; i < 25; i++) W[i] = -J[i];
        for (int i = 0; i < 5; i++) W[i * 5 + i] += dg;
        if (!lu5(W, piv)) return false;
        for (int s = 0; s < 8; s++) {
            const double* fs = f0;
            if (s > 0) {
                if (s <= 5) for (int i = 0; i < 5; i++) { double v = u[i]; for (int j = 0; j < s; j++) v += SB_BG_A[s][j] * ks[j][i]; U[i] = v; }
                else for (int i = 0; i < 5; i++) U[i] += ks[s - 1][i];
                sb_bg_rhs(U, P, du); fs = du;
            }
            for (int i = 0; i < 5; i++) { double v = fs[i]; for (int j = 0; j < s; j++) v += SB_BG_C[s][j] / dt * ks[j][i]; ks[s][i] = v; } // autonomous: dT = 0
            lusolve5(W, piv, ks[s]);
        }
        for (int i = 0; i < 5; i++) {
            unew[i] = U[i] + ks[7][i];
            for (int q = 0; q < 3; q++) { double s = 0; for (int j = 0; j < 8; j++) s += SB_BG_H[q][j] * ks[j][i]; K[q][i] = s; }
        }
        return true;
    }
    SB_HD void interp(const double* u0, const double* u1, double th, double* out) const {
        for (int i = 0; i < 5; i++) out[i] = (1 - th) * u0[i] + th * (u1[i] + (1 - th) * (K[0][i] + th * (K[1][i] + th * K[2][i])));
    }
    SB_HD void dinterp(const double* u0, const double* u1, double th, double dt, double* out) const {
        for (int i = 0; i < 5; i++) out[i] = (K[0][i] + th * (-2 * K[0][i] + 2 * K[1][i] + th * (-3 * K[1][i] + 3 * K[2][i] - 4 * th * K[2][i])) - u0[i] + u1[i]) / dt;
    }
};
SB_HD static inline double errnorm(const double* k8, const double* u0, const double* u1, double abstol, double reltol) {
    double s = 0;
    for (int i = 0; i < 5; i++) { double r = k8[i] / (abstol + reltol * fmax(fabs(u0[i]), fabs(u1[i]))); s += r * r; }
    return sqrt(s / 5);
}
SB_HD static inline double hubble_residual(const double* P, double aa, double tini) { double yy[5] = {aa, 0, 1, 1, 0}, g[5]; sb_bg_rhs(yy, P, g); return g[0] * tini / aa - 1.0; }
} // namespace bgsolve

// PI step controller with OrdinaryDiffEq's defaults for an order-5 method (β1 = 7/50, β2 = 2/25, γ = 0.9, q ∈ [1/5, 10], qoldinit 1e-4)
struct SbController {
    double qold, q11;
    SB_HD void init() { qold = 1e-4; q11 = 1.0; }
    SB_HD double q_of(double EEst) {
        if (EEst == 0.0) return 0.1;
        q11 = pow(EEst, 7.0 / 50.0);
        double q = q11 / pow(qold, 2.0 / 25.0);
        return fmax(0.1, fmin(5.0, q / 0.9));
    }
    SB_HD double accept(double dt, double q, double EEst) { qold = fmax(EEst, 1e-4); return dt / q; }
    SB_HD double reject(double dt) { return dt / fmin(5.0, q11 / 0.9); }
};

// The same PI controller for the perturbation kernels, in logarithms: q = EEst^(7/50) / qold^(2/25) = exp2((7/50) log2 EEst − (2/25) log2 qold) with log2 qold
// carried from the previous accepted step -- one log2 and one exp2 per attempt instead of two pow() (≈2 k cycles of a lone warp, 8 % of an attempt).
// Shared by every integrator kernel (explicit fma: the same bits everywhere).
#ifdef __CUDACC__
__device__ __forceinline__ double sb_ctl_q(double l, double lqold) { return exp2(fma(7.0 / 50.0, l, -(2.0 / 25.0) * lqold)); }   // q11 / qold^β2 from l = log2 EEst
__device__ __forceinline__ double sb_ctl_q11(double l) { return exp2((7.0 / 50.0) * l); }                                        // EEst^β1 (rejected steps)
__device__ __forceinline__ double sb_ctl_lq0() { double q0 = 1e-4; asm volatile("" : "+d"(q0)); return log2(q0); }               // log2 of the initial / smallest qold, by the device routine
#endif
namespace bgsolve {
// The background solve proper.  Knots (t, y, y') are written straight into the caller's arrays (capacity cap); returns the number
// of knots, or -1 if cap is too small / the solve failed before its first accepted step.
// info[0..5] = tau0, kappa0, taurec, retcode, naccept, nreject.
// tfix/nfix/dtlast (optional): LOCKSTEP mode for parameter lanes (sbm_solvebg_lock) -- take exactly the steps of another solve
// (its knots tfix[0..nfix-1]; the event step, in which a crosses 1, with length dtlast) without error control, so that the result
// is a smooth function of the parameters (the discrete map with frozen steps is what forward-mode AD differentiates as well).
// info[6] = length of the event step, info[7] = 1 if the knot count differs from nfix (lockstep lost).
SB_HD static inline int solve(const double* P, double tini, double tmax, double reltol, double abstol, int cap, double* T, double* Y, double* DY, double* info,
                              const double* tfix = nullptr, int nfix = 0, double dtlast = 0.0) {
    // a(τini) from ℋ = 1/τ (reference src/models/cosmologies.jl:76), Newton on ȧ τ / a − 1
    double a = sqrt(P[3] + P[4]) * tini;
    for (int it = 0; it < 100; it++) {
        double f = hubble_residual(P, a, tini), h = a * 1e-7, fp = (hubble_residual(P, a + h, tini) - hubble_residual(P, a - h, tini)) / (2 * h), an = a - f / fp;
        bool done = fabs(an - a) <= 1e-15 * fabs(a);
        a = an;
        if (done) break;
    }
    double u[5] = {a, 0, 1, 1, 0};
    int nT = 0, nD = 0; // knots with (t, y) written / with y' written (the first y' comes from the first accepted step)
    bool over = false;
#define SB_BG_PUSH_TY(tt_, u_) do { if (nT < cap) { T[nT] = (tt_); for (int i_ = 0; i_ < 5; i_++) Y[5 * nT + i_] = (u_)[i_]; nT++; } else over = true; } while (0)
#define SB_BG_PUSH_D(d_) do { if (nD < cap) { for (int i_ = 0; i_ < 5; i_++) DY[5 * nD + i_] = (d_)[i_]; nD++; } else over = true; } while (0)
    SB_BG_PUSH_TY(tini, u);
    Step S; SbController ctl; ctl.init();
    // automatic initial step (Hairer), as in OrdinaryDiffEq
    double dt;
    {
        double f0[5], f1[5], u1[5], sk[5], d0 = 0, d1 = 0, d2 = 0;
        sb_bg_rhs(u, P, f0);
        for (int i = 0; i < 5; i++) { sk[i] = abstol + fabs(u[i]) * reltol; d0 += (u[i] / sk[i]) * (u[i] / sk[i]); d1 += (f0[i] / sk[i]) * (f0[i] / sk[i]); }
        d0 = sqrt(d0 / 5); d1 = sqrt(d1 / 5);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        dt0 = fmin(dt0, tmax - tini);
        for (int i = 0; i < 5; i++) u1[i] = u[i] + dt0 * f0[i];
        sb_bg_rhs(u1, P, f1);
        for (int i = 0; i < 5; i++) { double r = (f1[i] - f0[i]) / sk[i]; d2 += r * r; }
        d2 = sqrt(d2 / 5) / dt0;
        double dm = fmax(d1, d2);
        double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2 + log10(dm)) / 5.0);
        dt = fmin(fmin(100 * dt0, dt1), tmax - tini);
    }
    double tt = tini; int rc = SB_RC_SUCCESS; long nacc = 0, nrej = 0; bool first = true;
    double tau0 = 0, kappa0 = 0, dtevent = 0;
    for (long it = 0;; it++) {
        if (it >= 100000) { rc = SB_RC_MAXITERS; break; }
        if (over) break;
        if (tfix) dt = (nacc + 1 <= nfix - 2) ? tfix[nacc + 1] - tt : dtlast; // prescribed steps; from the event step on: its length
        if (tt + dt > tmax) dt = tmax - tt;
        if (!S.run(P, u, dt)) { rc = SB_RC_UNSTABLE; break; }
        double EEst = errnorm(S.ks[7], u, S.unew, abstol, reltol);
        if (!isfinite(EEst)) { if (tfix) { rc = SB_RC_UNSTABLE; break; } nrej++; dt /= 5; if (dt < 1e-14 * tt) { rc = SB_RC_UNSTABLE; break; } continue; }
        double q = ctl.q_of(EEst);
        if (EEst > 1 && !tfix) { nrej++; dt = ctl.reject(dt); if (dt < 1e-14 * tt) { rc = SB_RC_DTMIN; break; } continue; }
        nacc++;
        double dtnew = ctl.accept(dt, q, EEst), d[5];
        if (first) { S.dinterp(u, S.unew, 0.0, dt, d); SB_BG_PUSH_D(d); first = false; }
        if (S.unew[0] - 1.0 >= 0) { // "today" callback: right-root of a − 1 on the dense output (src/solve.jl:158-202)
            double lo = 0, hi = 1, tmp[5];
            for (int b = 0; b < 200; b++) { double mid = 0.5 * (lo + hi); if (mid == lo || mid == hi) break; S.interp(u, S.unew, mid, tmp); if (tmp[0] - 1.0 >= 0) hi = mid; else lo = mid; }
            double uend[5], dend[5]; S.interp(u, S.unew, hi, uend);
            double dtr = hi * dt, tr = tt + dtr;
            // dense-output vectors of the shortened step (what OrdinaryDiffEq recomputes after moving t); S is not needed any more
            if (S.run(P, u, dtr)) S.dinterp(u, uend, 1.0, dtr, dend); else sb_bg_rhs(uend, P, dend);
            SB_BG_PUSH_TY(tr, uend); SB_BG_PUSH_D(dend);
            tau0 = tr; kappa0 = uend[1]; dtevent = dt;
            break;
        }
        S.dinterp(u, S.unew, 1.0, dt, d);
        tt = (tfix && nacc <= nfix - 2) ? tfix[nacc] : tt + dt; // lockstep: land on the prescribed knot bit for bit
        for (int i = 0; i < 5; i++) u[i] = S.unew[i];
        SB_BG_PUSH_TY(tt, u); SB_BG_PUSH_D(d);
        if (tt >= tmax) { tau0 = tt; kappa0 = u[1]; break; }
        dt = dtnew;
    }
#undef SB_BG_PUSH_TY
#undef SB_BG_PUSH_D
    const int nb = nT;
    // τrec: knot of maximal visibility over solver steps (src/solve.jl:183-189)
    double vmax = -1, taurec = tau0;
    for (int i = 0; i < nb; i++) { double g[5]; sb_bg_rhs(&Y[5 * i], P, g); double v = -g[1] * exp(-(Y[5 * i + 1] - kappa0)); if (v > vmax) { vmax = v; taurec = T[i]; } }
    info[0] = tau0; info[1] = kappa0; info[2] = taurec; info[3] = rc; info[4] = (double)nacc; info[5] = (double)nrej;
    info[6] = dtevent; info[7] = (tfix && nb != nfix) ? 1.0 : 0.0;
    if (over || nD != nb) return -1;
    return nb;
}
} // namespace bgsolve

extern "C" {

// out[0..15] = N, NPAR, NBETA, NB, LMAX, NX, W0WA, NNZ_FULL, FLOPS_F, FLOPS_LU, FLOPS_SOLVE, NLEVELS, NBLOCKS, P_KAPPA0, P_TAU0, NSLOT
int sbm_info(int* out) {
    int v[16] = {SB_N, SB_NPAR, SB_NBETA, SB_NB, SB_LMAX, SB_NX, SB_W0WA, SB_NNZ_FULL, SB_FLOPS_F, SB_FLOPS_LU, SB_FLOPS_SOLVE, SB_NLEVELS, SB_NBLOCKS, SB_P_KAPPA0, SB_P_TAU0, SB_NSLOT};
    memcpy(out, v, sizeof(v));
    return 0;
}
const char* sbm_key(void) { return SB_MODEL_KEY; }

// Background solve on the host.  Returns the number of spline knots, or -1 if cap is too small.
// info[0..7] = tau0, kappa0, taurec, retcode, naccept, nreject, length of the event step, 0.
int sbm_solvebg(const double* P, double tini, double tmax, double reltol, double abstol, int cap, double* t, double* y, double* dy, double* info) {
    return bgsolve::solve(P, tini, tmax, reltol, abstol, cap, t, y, dy, info);
}
// The same solve in LOCKSTEP with another one (parameter lanes of the sensitivity path): takes the prescribed knots tfix[nfix] of the
// primal solve and its event-step length dtlast = info[6] instead of choosing steps.  info[7] = 1 if the knot count came out different.
int sbm_solvebg_lock(const double* P, double tini, double tmax, int nfix, const double* tfix, double dtlast, int cap, double* t, double* y, double* dy, double* info) {
    if (nfix < 3 || !tfix || !(dtlast > 0)) return -1;
    return bgsolve::solve(P, tini, tmax, 1.0, 1.0, cap, t, y, dy, info, tfix, nfix, dtlast);
}

} // extern "C"

// ================================================================================================ DEVICE
// Batched background solve: one thread per cosmology.  A parameter sweep has hundreds to thousands of
// independent 5-unknown stiff solves of ≈10³ steps each; the threads of a warp run the same step loop (divergence only in the
// accept/reject tails and the step counts), the working set (stage vectors, 5×5 factors) lives in registers/local memory.
// P[c] gets kappa0 and tau0 filled in (callback semantics, src/solve.jl:183-189) so that the perturbation launch can follow on
// the same stream without a host round trip.
__global__ void __launch_bounds__(32) sb_solvebg_kernel(int n, double* __restrict__ P, double tini, double tmax, double reltol, double abstol, int cap, double* __restrict__ t, double* __restrict__ y,
                                                         double* __restrict__ dy, double* __restrict__ info, int* __restrict__ nb) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double* Pc = P + (size_t)c * SB_NPAR;
    double inf[8];
    const int r = bgsolve::solve(Pc, tini, tmax, reltol, abstol, cap, t + (size_t)c * cap, y + (size_t)c * cap * 5, dy + (size_t)c * cap * 5, inf);
    nb[c] = r;
    for (int j = 0; j < 8; j++) info[(size_t)c * 8 + j] = inf[j];
    Pc[SB_P_KAPPA0] = inf[1]; Pc[SB_P_TAU0] = inf[0];
}

// β-table.  The background spline is piecewise cubic on the solver's knots, so β_m(y(τ)) is smooth only *between* knots:
// every knot interval [t_j, t_{j+1}] is subdivided uniformly into msub sub-intervals; node n = j·msub + s sits at
// τ = t_j + s·(t_{j+1} − t_j)/msub and stores β_m and dβ_m/dτ (exact, chain rule through the spline derivative).
// Cubic Hermite in τ between nodes.  The table replaces per-stage re-evaluation of the thermodynamics (exp/tanh/pow chains)
// inside the integrator; its interpolation error (≈1e-12 relative, tests/test_gpu_parity.py) is far below the solver tolerance.
__global__ void sb_table_kernel(const double* __restrict__ P, SbSpline spl, int msub, double* __restrict__ tab) {
    const int nnode = (spl.nb - 1) * msub + 1;
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    int j = n / msub, sidx = n % msub;
    if (j > spl.nb - 2) { j = spl.nb - 2; sidx = msub; }
    const double h = spl.t[j + 1] - spl.t[j];
    const double tau = (sidx == msub) ? spl.t[j + 1] : spl.t[j] + sidx * (h / msub);
    double y[5], yp[5], beta[SB_NBETA], betad[SB_NBETA];
    const double d0 = tau - spl.t[j], d1 = tau - spl.t[j + 1];
    for (int q = 0; q < 5; q++) {
        double u0 = spl.y[5 * j + q], u1 = spl.y[5 * j + 5 + q], v0 = spl.dy[5 * j + q], v1 = spl.dy[5 * j + 5 + q];
        double c1 = (u1 - u0 - v0 * h) / (h * h), c2 = (v1 - v0 - 2 * c1 * h) / (h * h);
        y[q] = u0 + d0 * v0 + d0 * d0 * (c1 + d1 * c2);
        yp[q] = v0 + 2 * d0 * (c1 + d1 * c2) + d0 * d0 * c2;
    }
    sb_beta(tau, y, yp, P, beta, betad);
    double* o = tab + (size_t)n * 2 * SB_NBETA;
    for (int m = 0; m < SB_NBETA; m++) { o[m] = beta[m]; o[SB_NBETA + m] = betad[m]; }
}

struct SbTable {
    int nb, msub, nlut;      // knots, sub-intervals per knot interval, size of the interval look-up table
    double s0, inv_dsl;      // lut[q] = knot interval containing exp(s0 + q/inv_dsl)
    const double* t;         // knots
    const int* lut;
    const double* tab;       // [(nb-1)·msub + 1][2][NBETA]
};

// One background cosmology as the integrator sees it.  Layout = sbm_cosmo_t of include/symboltz_b200.h (112 bytes).
struct SbCosmo {
    const double* P;
    SbSpline spl;
    SbTable tb;
    double tend;
    const double* saveat; // [nsave] save times (may be null when nsave = 0)
    const double* srcbg;  // [nsave][SB_SRCBG_STRIDE] per-save-time background of the fused source evaluation (sbm_srcbg), or null
    double taurec;        // time of maximal visibility (lensing source Sψ)
};
static_assert(sizeof(SbCosmo) == 128, "SbCosmo must match sbm_cosmo_t");
// Fused source evaluation request (= sbm_src_t of include/symboltz_b200.h): host struct passed by pointer to the *_src entry points.
struct sbm_src_t {
    const double* dsrcbg; // device [nsave][sbm_srcbg_stride()] from sbm_srcbg (single-cosmology calls; batched calls take it from sbm_cosmo_t)
    double* dS;           // device [nk][nS][nsave]
    int nS, scale_k;      // nS = 2 (ST, SE) or 3 (+ Sψ); scale_k != 0: (k·ST, k²·SE)
    double taurec;        // single-cosmology calls
};
#define SB_COSMO_DOUBLES 16

struct SbSolveArgs {
    SbCosmo c0;            // the cosmology of a single-cosmology launch
    const SbCosmo* cosmos; // batched launch (sbm_solvept_batch): device array of cosmologies and the cosmology index of each mode
    const int* cosmo_of;
    int nk;
    const double *ks, *tini;
    const int* order;
    int nsave;
    double reltol, abstol;
    int maxiters;
    double *usave, *uend;
    int* retcode;
    long long* stats; // [nk][4] = naccept, nreject, nf, nsolve
    int* queue;
    double* trace;    // optional debug trace of mode 0: (t, dt, EEst) per attempted step
    int ntrace;
    // optional static schedule (sbm_solvept_sched): warp w runs items [ibeg[w], ibeg[w+1]); item = (mode, quota, cont)
    const int* items;
    const int* ibeg;
    double* cont;     // [nk][SB_CONT] continuation records
    int* flags;       // [nk]: 0 not yet parked, 1 parked (record valid), 2 finished inside the first piece
    // fused source evaluation (sbm_solvept_src & co.): S[nk][nS][nsave] formed from the dense output at every save time
    double* S;
    int nS, scale_k;
    // parameter lanes in lockstep (sbm_solvept_lanes): G warps of one CTA integrate the same mode for G cosmologies with ONE step controller
    int G;
    double tendc;       // common end of the lockstep phase = the smallest end time of the lanes; each lane closes with a private step to its own
    double invdelta[8]; // 1/δ_j of lane j (δ_j = parameter step of the lane in the units of the wanted derivative); [0] unused
};
#define SB_GROUP_MAX 8
#define SB_XCH_HDR 8 // doubles at the head of the CTA's exchange area: [0] work index, [1] initial step
#define SB_CONT (SB_N + 12)

// shared-memory layout per warp (doubles)
#define SB_SM_U 0
#define SB_SM_UU (SB_SM_U + SB_N)
#define SB_SM_K (SB_SM_UU + SB_N)
#define SB_SM_DT (SB_SM_K + 8 * SB_N)
#define SB_SM_F0 (SB_SM_DT + SB_N)
#define SB_SM_DI (SB_SM_F0 + SB_N)
#define SB_SM_UP (SB_SM_DI + SB_N)
#define SB_SM_MM (SB_SM_UP + SB_N)
#define SB_SM_BLK (SB_SM_MM + SB_N) // must follow mm: the factorisation scatters into mm and blk through one index
#define SB_SM_ZP (SB_SM_BLK + SB_TOPSTORE)
#define SB_SM_ZQ (SB_SM_ZP + SB_N)
#define SB_SM_BS (SB_SM_ZQ + SB_N)
#define SB_SM_BD (SB_SM_BS + 6 * SB_NB)
#define SB_SM_KP (SB_SM_BD + SB_NB)
#define SB_SM_COSMO (SB_SM_KP + 8 + 48)  // batched launches: the current mode's SbCosmo
#define SB_SM_SBUF (SB_SM_COSMO + 16)    // fused sources: [3][SB_SWIN] values of the current window of save times
#define SB_SWIN 16
// SB_TMA (north star: "tables ... staged to shared memory with TMA"): the 2 x 2 β-table rows of the six stage times of an attempt are
// staged into shared memory by six 1.3 KB bulk async copies (cp.async.bulk + mbarrier, SASS UBLKCP) instead of being read through L1
// with prefetch + __ldg.  Same-box A/B (profiles/integrate_r2.md, r2d): +1.8 % attempts/s on a saturated launch, −1.6 % latency of a
// lone warp, neutral at the bench size; on.
#ifndef SB_TMA
#define SB_TMA 1
#endif
#if SB_TMA
#define SB_SM_TAB ((SB_SM_SBUF + 3 * SB_SWIN + 1) & ~1) // 16-byte aligned: [6 slots][4 rows][NBETA]
#define SB_SM_MBAR (SB_SM_TAB + 6 * 4 * SB_NBETA)
#define SB_SM_BD5 (SB_SM_MBAR + 2) // derivative basis at the last stage time: the next attempt's slot 0 (SB_DEFER)
#define SB_SM_DOUBLES ((SB_SM_BD5 + SB_NB + 1) & ~1)
#else
#define SB_SM_DOUBLES ((SB_SM_SBUF + 3 * SB_SWIN + 1) & ~1)
#endif
#define SB_SM_BYTES_WARP (SB_SM_DOUBLES * 8)
#define SB_WARPS_PER_CTA 1
#ifndef SB_MINBLOCKS
#define SB_MINBLOCKS 8
#endif
// Warps per SM come in steps of four (each of the four sub-partitions holds 16 K registers): 8 warps up to 255 registers, 12 up
// to 168.  Measured (profiles/integrate_r1.md): 12 warps with 168 registers (150-360 B of spills) gain 4 % on a long uniform
// queue but lose 30 % of single-mode latency, and the config-4 sweep (chunked launches with tails) is 8-14 % SLOWER -- so both
// instantiations run 8 warps with all the registers.  The batched instantiation keeps its own knob.
#ifndef SB_MINBLOCKS_BATCH
#define SB_MINBLOCKS_BATCH 8
#endif
#define SB_SM_BYTES (SB_SM_BYTES_WARP * SB_WARPS_PER_CTA)
#define SB_NBR ((SB_NB + 31) / 32)
#ifndef SB_TOPFAST
#define SB_TOPFAST 1
#endif
#ifndef SB_GJ_REDUX
#define SB_GJ_REDUX 1
#endif
#ifndef SB_GJ_INPLACE
#define SB_GJ_INPLACE 1
#endif
#ifndef SB_Z3
#define SB_Z3 1
#endif
// Division-free pivots along the phase-0 paths (see sb_factor); the sequential form is kept for long paths (register arrays of SB_PL entries)
#ifndef SB_PATH_CONT
#define SB_PATH_CONT (SB_PL <= 12)
#endif
// Stage loop of the warp-per-mode kernel: absent tableau terms as exact zeros instead of (warp-uniform) branches / the loop unrolled (A/B: profiles/integrate_r2.md)
#ifndef SB_STAGE_BF
#define SB_STAGE_BF 0
#endif
#ifndef SB_STAGE_UNROLL
#define SB_STAGE_UNROLL 0 // unrolled: +2.5 % attempts/s on a saturated launch WITHOUT save times, but with the 300 save times of the C_l workload the step goes from 57.5 to 84 ms (the source evaluation's code no longer stays in the instruction cache next to a 7x larger stage body); the select form SB_STAGE_BF: -0.2 %
#endif
#ifndef SB_SPLIT_UNROLL
#define SB_SPLIT_UNROLL 0
#endif
// Table rows of the stage times requested at the start of the attempt and swept right before the first stage evaluation (see sb_split_basis_issue)
#ifndef SB_DEFER
#define SB_DEFER (SB_TMA && SB_BSLOT)
#endif
#ifndef SB_KEEP_MAX
#define SB_KEEP_MAX 24 // doubles per lane that sb_bsolve may hold across its top step (see KEEP there)
#endif
// packed schedule fields (generator: lower.py).  SB_IDXBITS = 8 when N <= 254 and NB <= 255 (byte extraction), else 10 (N <= 1022).
#if SB_IDXBITS == 8
#define SB_E_COL(ix) ((ix) & 255u)
#define SB_E_B(ix) (((ix) >> 8) & 255u)
#define SB_E_KIND(ix) (((ix) >> 16) & 15u)
#define SB_E_TGT(ix) ((ix) >> 20)
#define SB_LO16(v) ((v) & 255u)
#define SB_HI16(v) (((v) >> 8) & 255u)
#define SB_P_START(v) ((int)((v) & 255u))
#define SB_P_LEN(v) ((int)(((v) >> 8) & 255u))
#define SB_P_PAR(v) ((int)(((v) >> 16) & 255u))
#define SB_KID(v, c) ((int)(((v) >> (8 * (c))) & 255u))
#define SB_NOKID 255
#define SB_NOPAR 255
#else
#define SB_E_COL(ix) ((ix) & 1023u)
#define SB_E_B(ix) (((ix) >> 10) & 1023u)
#define SB_E_KIND(ix) (((ix) >> 20) & 3u)
#define SB_E_TGT(ix) ((ix) >> 22)
#define SB_LO16(v) ((v) & 0xFFFFu)
#define SB_HI16(v) ((v) >> 16)
#define SB_P_START(v) ((int)((v) & 4095u))
#define SB_P_LEN(v) ((int)(((v) >> 12) & 255u))
#define SB_P_PAR(v) ((int)((v) >> 20))
#define SB_KID(v, c) ((int)(((v) >> (10 * (c))) & 1023u))
#define SB_NOKID 1023
#define SB_NOPAR 4095
#endif

// SB_RCP: 1/x of the pivots (13 per attempt, all on the dependent chain of the factorisation) as MUFU.RCP64H seed + two Newton steps
// instead of the IEEE division sequence with its slow-path branch.  Same-box A/B (r2d): +4 % attempts/s, −3.6 % lone-warp latency; on.
#ifndef SB_RCP
#define SB_RCP 1
#endif
#ifndef SB_BSLOT   // slot-major basis sweep (TMA build only), see sb_basis_batch
#define SB_BSLOT 1
#endif
#ifndef SB_NORMRCP // error norm: k8 · rcp(scale) instead of k8 / scale
#define SB_NORMRCP 1
#endif
__device__ __forceinline__ double sb_rcp(double x) {
#if SB_RCP
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return 1.0 / x;
#endif
}
// Scalar pieces of the attempt's tail, shared by the kernels, with multiplications by reciprocals where the reference formula divides (each IEEE division is a
// ~200-cycle dependent chain on a lone warp; the constants differ from the quotients by at most an ulp)
__device__ __forceinline__ double sb_ctl_qacc(double r) { return fmax(0.1, fmin(5.0, r * (1.0 / 0.9))); }   // clamp(q / γ_c) of an accepted step
__device__ __forceinline__ double sb_ctl_qrej(double r11) { return fmin(5.0, r11 * (1.0 / 0.9)); }           // min(1/qmin, q11 / γ_c) of a rejected step
__device__ __forceinline__ double sb_rms(double sumsq) { return sqrt(sumsq * (1.0 / SB_N)); }                // RMS norm from the sum of squares
__device__ __forceinline__ double sb_wdiag(double dt) { return sb_rcp(SB_R5_GAMMA * dt); }                   // 1 / (γ dt)
__device__ __forceinline__ unsigned sb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SB_FULL, v, o);
    return v;
}

// Row windows.  Row i of the state belongs to lane i % 32 and round i / 32.  The warp-per-mode kernels loop over all SB_R rounds; in the
// split kernel (SPL) each warp of the CTA owns one round, rb, held in the round-0 slots of its SbLane (see SbLane::load_split).
#define SB_WMAX SB_WDR(0)                                // rounds are ordered by decreasing row length: round 0 is the widest
#define SB_NRL (SPL ? 1 : SB_R)                          // rounds this warp loops over
#define SB_WDL(r) (SPL ? SB_WMAX : SB_WDR(r))            // ELL width of its round r
#define SB_EOL(r) (SPL ? 0 : SB_EOFF(r))                 // first ELL slot of its round r
#define SB_ROW(r) ((rb + (r)) * 32 + lane)               // state index of its row in round r
// Per-lane schedules produced by the code generator, kept in registers for the lifetime of the kernel
// (every loop over them is fully unrolled so that the arrays never become addressable local memory).
struct SbLane {
    double ec[SB_ELLN];         // J_local ELL coefficients of the rows owned by this lane
    unsigned ei[SB_ELLN]; // col | basis<<8 | role<<16 | target<<20
    double pqc[SB_R * 2];      // hub vectors p, q at the owned rows
    unsigned pqi[SB_R];
    double phc[SB_TPH], psc[SB_TPS]; // hub functionals φ, ψ: one term per lane and slot
    unsigned phi[SB_TPH], psi[SB_TPS];
    unsigned ph[SB_PR], pk[SB_PR]; // owned phase-0 elimination paths (contiguous index ranges): start|len<<8|parent<<16, first-vertex children
    unsigned rh[SB_TR + 1], rk[SB_TR + 1]; // owned multi-vertex root paths (the part of the top that is neither in a block nor a single vertex)
    unsigned tv, dk;           // top vertex owned by this lane (block member: lane = block*8+row, or single root) and its forest children
    unsigned dkz;              // dk with absent children replaced by a top-block vertex (its multiplier mm is always 0): unconditional gathers
    unsigned bp[SB_NBR];       // basis m -> beta index | (kpow+3)<<8
    // Split mode (sb_integrate_split_kernel): the warp owns ONE round of rows, rb; its ELL entries and hub-vector entries are loaded into
    // the slots of "round 0" (width SB_WMAX, zero coefficients beyond the round's own width), so that the row loops below run over a single
    // round with compile-time register indices.  Everything that is per lane, not per row (paths, top vertices, hub functionals, basis
    // descriptors), is the same in every warp.
    __device__ __forceinline__ void load_split(int lane, int rb) {
        load(lane);
        double ec0[SB_WMAX]; unsigned ei0[SB_WMAX]; double p0 = 0, q0 = 0; unsigned pq0 = 0;
#pragma unroll
        for (int r = 0; r < SB_R; r++) {
            if (r == rb) {
#pragma unroll
                for (int w = 0; w < SB_WMAX; w++) { ec0[w] = (w < SB_WDR(r)) ? ec[SB_EOFF(r) + (w < SB_WDR(r) ? w : 0)] : 0.0; ei0[w] = (w < SB_WDR(r)) ? ei[SB_EOFF(r) + (w < SB_WDR(r) ? w : 0)] : 0u; }
                p0 = pqc[2 * r]; q0 = pqc[2 * r + 1]; pq0 = pqi[r];
            }
        }
#pragma unroll
        for (int w = 0; w < SB_WMAX; w++) { ec[w] = ec0[w]; ei[w] = ei0[w]; }
        pqc[0] = p0; pqc[1] = q0; pqi[0] = pq0;
    }
    __device__ __forceinline__ void load(int lane) {
#pragma unroll
        for (int e = 0; e < SB_ELLN; e++) { ec[e] = sb_ell_coef[e * 32 + lane]; ei[e] = sb_ell_idx[e * 32 + lane]; }
#pragma unroll
        for (int r = 0; r < SB_R; r++) { pqc[2 * r] = sb_pq_coef[(2 * r) * 32 + lane]; pqc[2 * r + 1] = sb_pq_coef[(2 * r + 1) * 32 + lane]; pqi[r] = sb_pq_idx[r * 32 + lane]; }
#pragma unroll
        for (int t = 0; t < SB_TPH; t++) { phc[t] = sb_phi_coef[t * 32 + lane]; phi[t] = sb_phi_idx[t * 32 + lane]; }
#pragma unroll
        for (int t = 0; t < SB_TPS; t++) { psc[t] = sb_psi_coef[t * 32 + lane]; psi[t] = sb_psi_idx[t * 32 + lane]; }
#pragma unroll
        for (int q = 0; q < SB_PR; q++) { ph[q] = sb_path_head[q * 32 + lane]; pk[q] = sb_path_kids[q * 32 + lane]; }
#pragma unroll
        for (int q = 0; q < SB_TR; q++) { rh[q] = sb_root_head[q * 32 + lane]; rk[q] = sb_root_kids[q * 32 + lane]; }
        dk = sb_top_kids[lane]; tv = sb_top_vert[lane];
        {
            const unsigned z = sb_top_vert[0];
            dkz = 0;
#pragma unroll
            for (int c = 0; c < 3; c++) { const unsigned kid = (unsigned)SB_KID(dk, c); dkz |= (kid == SB_NOKID ? z : kid) << (SB_IDXBITS == 8 ? 8 * c : 10 * c); }
        }
#pragma unroll
        for (int r = 0; r < SB_NBR; r++) { int m = r * 32 + lane; bp[r] = (m < SB_NB) ? sb_basis_pack[m] : 0u; }
    }
};

// basis functions b_m = k^e β_m(τ) (and optionally ḃ_m) from the β-table; lanes over m
__device__ __forceinline__ int sb_interval(const SbTable& tb, double tau) { // knot interval containing tau (log look-up + scan)
    int q = (int)((log(tau) - tb.s0) * tb.inv_dsl);
    q = max(0, min(q, tb.nlut - 1));
    int j = __ldg(tb.lut + q);
    while (j < tb.nb - 2 && __ldg(tb.t + j + 1) <= tau) j++;
    return j;
}
// jhint: an interval index known to be <= the interval of tau (times only move forward inside a mode)
__device__ __forceinline__ int sb_basis_at(const SbLane& S, const SbTable& tb, double tau, int jhint, const double* kp /*smem: k^-3..k^3 at [0..6]*/, double* b, double* bd, int lane) {
    int j = jhint;
    while (j < tb.nb - 2 && __ldg(tb.t + j + 1) <= tau) j++;
    const double tj = __ldg(tb.t + j), hj = __ldg(tb.t + j + 1) - tj;
    const double hs = hj / tb.msub;
    double f = (tau - tj) / hs;
    int sidx = max(0, min((int)f, tb.msub - 1));
    const double w = f - sidx, w1 = w - 1.0;
    const double h00 = (1 + 2 * w) * w1 * w1, h10 = w * w1 * w1 * hs, h01 = w * w * (3 - 2 * w), h11 = w * w * w1 * hs;
    const double* n0 = tb.tab + ((size_t)j * tb.msub + sidx) * 2 * SB_NBETA;
    const double* n1 = n0 + 2 * SB_NBETA;
#pragma unroll
    for (int r = 0; r < SB_NBR; r++) {
        int m = r * 32 + lane;
        if (m < SB_NB) {
            int be = SB_LO16(S.bp[r]);
            double v0 = __ldg(n0 + be), d0 = __ldg(n0 + SB_NBETA + be), v1 = __ldg(n1 + be), d1 = __ldg(n1 + SB_NBETA + be);
            double kk = kp[SB_HI16(S.bp[r])];
            b[m] = kk * (h00 * v0 + h10 * d0 + h01 * v1 + h11 * d1);
            if (bd) bd[m] = kk * (6 * w * w1 * (v0 - v1) / hs + (3 * w - 1) * w1 * d0 + w * (3 * w - 2) * d1); // derivative of the Hermite cubic
        }
    }
    return j;
}

// Batched look-up of the basis at the stage times of one Rosenbrock attempt: slots 1..5 = t + c_s·dt (and, if with0, slot 0 = t
// together with ḃ(t)).  Lanes 0..5 locate their slot's table node in parallel and publish (node, w, hs, τ) through shared memory;
// then all lanes sweep the (slot, basis) items, so that every L2 access of the attempt is in flight at once.
// Locate stage time tau in the table (search forward from knot interval j) and put the table node and the cubic-Hermite weights of
// (v0, d0, v1, d1), plus the derivative weights, into sp[0..7]; returns the knot interval.
__device__ __forceinline__ int sb_slot_locate(const SbTable& tb, double tau, int j, double* sp) {
    while (j < tb.nb - 2 && __ldg(tb.t + j + 1) <= tau) j++;
    // (Newton reciprocals instead of three dependent IEEE divisions: the search sits at the head of every attempt)
    const double tj = __ldg(tb.t + j), dtj = __ldg(tb.t + j + 1) - tj;
    const double hs = dtj * sb_rcp((double)tb.msub), ihs = (double)tb.msub * sb_rcp(dtj);
    const double f = (tau - tj) * ihs;
    const int sidx = max(0, min((int)f, tb.msub - 1));
    const double w = f - sidx, w1 = w - 1.0;
    sp[0] = __longlong_as_double((long long)j * tb.msub + sidx);
    sp[1] = (1 + 2 * w) * w1 * w1; sp[2] = w * w1 * w1 * hs; sp[3] = w * w * (3 - 2 * w); sp[4] = w * w * w1 * hs;
    sp[5] = 6 * w * w1 * ihs; sp[6] = (3 * w - 1) * w1; sp[7] = w * (3 * w - 2);
    return j;
}
__device__ __forceinline__ double sb_hermite_val(double kk, double w1, double v0, double w2, double d0, double w3, double v1, double w4, double d1) { return kk * (w1 * v0 + w2 * d0 + w3 * v1 + w4 * d1); }
__device__ __forceinline__ double sb_hermite_der(double kk, double s5, double v0, double v1, double s6, double d0, double s7, double d1) { return kk * (s5 * (v0 - v1) + s6 * d0 + s7 * d1); }
__device__ __forceinline__ int sb_basis_batch(const SbLane& S, const SbTable& tb, double t, double dt, int jt, bool with0, const double* kp, double* bs, double* bdv, double* slotp /*smem 6x8*/, int lane,
                                              double* tabs = nullptr, unsigned* parity = nullptr) {
    int jend = jt;
#if SB_TMA
    const int s0t = with0 ? 0 : 1;
    const unsigned mbar = sb_smem_u32(tabs + 6 * 4 * SB_NBETA);
    constexpr unsigned ROWB = 4 * SB_NBETA * 8; // two nodes x (β, dβ/dτ): contiguous in the table
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the previous attempt's generic-proxy reads of the staging area come first
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((6 - s0t) * ROWB) : "memory");
#endif
    if (lane < 6) {
        const double tau = fma((lane == 0) ? 0.0 : cc[lane], dt, t);
        double* sp = slotp + lane * 8;
        jend = sb_slot_locate(tb, tau, jt, sp);
#if SB_TMA
        if (lane >= s0t) {
            const double* src = tb.tab + (size_t)__double_as_longlong(sp[0]) * 2 * SB_NBETA;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb_smem_u32(tabs + lane * 4 * SB_NBETA)), "l"(src), "r"(ROWB), "r"(mbar) : "memory");
        }
#endif
    }
    jend = __shfl_sync(SB_FULL, jend, 5);
    __syncwarp();
    const int s0 = with0 ? 0 : 1;
#if SB_TMA
    {
        const unsigned ph = *parity;
        asm volatile("{\n .reg .pred p;\n SB_WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra SB_WAIT_%=;\n}" ::"r"(mbar), "r"(ph) : "memory");
        *parity = ph ^ 1u;
    }
#if SB_BSLOT
    // slot-major sweep on the lane's register-resident basis descriptors (S.bp): no integer division, no table load per item, the slot's
    // Hermite weights are warp-uniform broadcasts; same expression per item as the item-major sweep below (bit-identical)
    {
        double kk[SB_NBR];
        int be[SB_NBR];
#pragma unroll
        for (int r = 0; r < SB_NBR; r++) { be[r] = SB_LO16(S.bp[r]); kk[r] = kp[SB_HI16(S.bp[r])]; }
#pragma unroll
        for (int s = 0; s < 6; s++) {
            if (s == 0 && !with0) continue;
            const double* sp = slotp + s * 8;
            const double w1 = sp[1], w2 = sp[2], w3 = sp[3], w4 = sp[4];
            const double* n0 = tabs + s * 4 * SB_NBETA;
#pragma unroll
            for (int r = 0; r < SB_NBR; r++) {
                const int m = r * 32 + lane;
                if (m < SB_NB) {
                    const double v0 = n0[be[r]], d0 = n0[SB_NBETA + be[r]], v1 = n0[2 * SB_NBETA + be[r]], d1 = n0[3 * SB_NBETA + be[r]];
                    bs[s * SB_NB + m] = sb_hermite_val(kk[r], w1, v0, w2, d0, w3, v1, w4, d1);
                    if (s == 0) bdv[m] = sb_hermite_der(kk[r], sp[5], v0, v1, sp[6], d0, sp[7], d1);
                }
            }
        }
    }
#else
    for (int q = s0 * SB_NB + lane; q < 6 * SB_NB; q += SB_WARP) {
        const int s = q / SB_NB, m = q - s * SB_NB;
        const double* sp = slotp + s * 8;
        const unsigned pk = sb_basis_pack[m];
        const int be = SB_LO16(pk);
        const double* n0 = tabs + s * 4 * SB_NBETA;
        const double v0 = n0[be], d0 = n0[SB_NBETA + be], v1 = n0[2 * SB_NBETA + be], d1 = n0[3 * SB_NBETA + be];
        const double kk = kp[SB_HI16(pk)];
        bs[q] = kk * (sp[1] * v0 + sp[2] * d0 + sp[3] * v1 + sp[4] * d1);
        if (s == 0) bdv[m] = kk * (sp[5] * (v0 - v1) + sp[6] * d0 + sp[7] * d1);
    }
#endif
    return jend;
#endif
    // pull the 2 x 2 table rows of every slot into L1 first (one round trip to L2 for the whole attempt instead of one per sweep iteration)
    {
        constexpr int LINES = (4 * SB_NBETA * 8 + 127) / 128 + 1;
        for (int q = s0 * LINES + lane; q < 6 * LINES; q += SB_WARP) {
            const int s = q / LINES, l = q - s * LINES;
            const long long node = __double_as_longlong(slotp[s * 8]);
            const char* pa = (const char*)(tb.tab + (size_t)node * 2 * SB_NBETA) + l * 128;
            if (l * 128 < 4 * SB_NBETA * 8 + 120) asm volatile("prefetch.global.L1 [%0];" ::"l"(pa));
        }
    }
    for (int q = s0 * SB_NB + lane; q < 6 * SB_NB; q += SB_WARP) {
        const int s = q / SB_NB, m = q - s * SB_NB;
        const double* sp = slotp + s * 8;
        const long long node = __double_as_longlong(sp[0]);
        const unsigned pk = sb_basis_pack[m];
        const int be = SB_LO16(pk);
        const double* n0 = tb.tab + (size_t)node * 2 * SB_NBETA;
        const double v0 = __ldg(n0 + be), d0 = __ldg(n0 + SB_NBETA + be), v1 = __ldg(n0 + 2 * SB_NBETA + be), d1 = __ldg(n0 + 3 * SB_NBETA + be);
        const double kk = kp[SB_HI16(pk)];
        bs[q] = kk * (sp[1] * v0 + sp[2] * d0 + sp[3] * v1 + sp[4] * d1);
        if (s == 0) bdv[m] = kk * (sp[5] * (v0 - v1) + sp[6] * d0 + sp[7] * d1);
    }
    return jend;
}

// Barriers of the split kernel (sb_integrate_split_kernel): its SB_R row warps synchronise on named barrier 2; the controller warp, which
// only computes the step controller's powers, is not part of it (barriers 3 and 4 are its mailbox, barrier 1 joins the warps that sweep the basis).
#define SB_SPLIT_NT (SB_WARP * SB_R)
#define SB_SPLIT_THREADS (SB_SPLIT_NT + SB_WARP)
__device__ __forceinline__ void sb_rows_sync() { asm volatile("bar.sync 2, %0;" ::"n"(SB_SPLIT_NT) : "memory"); }
__device__ __forceinline__ bool sb_rows_or(bool p) {
    unsigned r;
    asm volatile("{\n .reg .pred p, q;\n setp.ne.u32 p, %1, 0;\n bar.red.or.pred q, 2, %2, p;\n selp.u32 %0, 1, 0, q;\n}" : "=r"(r) : "r"((unsigned)p), "n"(SB_SPLIT_NT) : "memory");
    return r != 0;
}
// out = J(b)·U : owned ELL rows + the two hub functionals Φ̇ = φᵀU, Ψ = ψᵀU.
// FUSE: out = J(b)·U + radd[r] + hd·dT  (the stage right-hand side is assembled in the same pass: one store per row)
template <bool FUSE, bool SPL = false>
__device__ __forceinline__ void sb_eval_f(const SbLane& S, const double* b, const double* U, double* out, int lane, const double* radd = nullptr, double hd = 0.0, const double* dT = nullptr,
                                          double* hubs = nullptr /* out: Φ̇ = φᵀU, Ψ = ψᵀU */, int rb = 0) {
    double sphi = 0, spsi = 0;
#pragma unroll
    for (int t = 0; t < SB_TPH; t++) sphi += S.phc[t] * b[SB_HI16(S.phi[t])] * U[SB_LO16(S.phi[t])];
#pragma unroll
    for (int t = 0; t < SB_TPS; t++) spsi += S.psc[t] * b[SB_HI16(S.psi[t])] * U[SB_LO16(S.psi[t])];
    // the butterfly reduction of the two hub sums is interleaved, level by level, with the (independent) ELL row sums so that
    // the shuffle latencies are covered by useful work; same operand order as warp_sum
    double acc[SB_NRL];
#pragma unroll
    for (int r = 0; r < SB_NRL; r++) {
        const int o = 16 >> r;
        double t1 = 0, t2 = 0;
        if (o > 0) { t1 = __shfl_xor_sync(SB_FULL, sphi, o); t2 = __shfl_xor_sync(SB_FULL, spsi, o); }
        double a_ = 0;
#pragma unroll
        for (int w = 0; w < SB_WDL(r); w++) { const int e = SB_EOL(r) + w; a_ += S.ec[e] * b[SB_E_B(S.ei[e])] * U[SB_E_COL(S.ei[e])]; }
        acc[r] = a_;
        if (o > 0) { sphi += t1; spsi += t2; }
    }
#pragma unroll
    for (int o = (SB_NRL < 5 ? 16 >> SB_NRL : 0); o > 0; o >>= 1) { sphi += __shfl_xor_sync(SB_FULL, sphi, o); spsi += __shfl_xor_sync(SB_FULL, spsi, o); }
#pragma unroll
    for (int r = 0; r < SB_NRL; r++) {
        const int i = SB_ROW(r);
        double a_ = acc[r] + S.pqc[2 * r] * b[SB_LO16(S.pqi[r])] * sphi + S.pqc[2 * r + 1] * b[SB_HI16(S.pqi[r])] * spsi;
        if (FUSE) { if (i < SB_N) a_ += radd[r] + hd * dT[i]; }
        if (i < SB_N) out[i] = a_;
    }
    if (hubs) { hubs[0] = sphi; hubs[1] = spsi; }
    if (!SPL) __syncwarp(); // split kernel: the caller follows with a CTA barrier
}
// ∂f/∂τ = J'u with J' = J_local(ḃ) + ṗ φᵀ + p φ̇ᵀ + q̇ ψᵀ + q ψ̇ᵀ
template <bool SPL = false>
__device__ __forceinline__ void sb_eval_dT(const SbLane& S, const double* b, const double* bd, const double* U, double* out, int lane, int rb = 0) {
    double sphi = 0, spsi = 0, sphid = 0, spsid = 0;
#pragma unroll
    for (int t = 0; t < SB_TPH; t++) { double cu = S.phc[t] * U[SB_LO16(S.phi[t])]; sphi += cu * b[SB_HI16(S.phi[t])]; sphid += cu * bd[SB_HI16(S.phi[t])]; }
#pragma unroll
    for (int t = 0; t < SB_TPS; t++) { double cu = S.psc[t] * U[SB_LO16(S.psi[t])]; spsi += cu * b[SB_HI16(S.psi[t])]; spsid += cu * bd[SB_HI16(S.psi[t])]; }
    sphi = warp_sum(sphi); spsi = warp_sum(spsi); sphid = warp_sum(sphid); spsid = warp_sum(spsid);
#pragma unroll
    for (int r = 0; r < SB_NRL; r++) {
        const int pb = SB_LO16(S.pqi[r]), qb = SB_HI16(S.pqi[r]);
        double acc = S.pqc[2 * r] * (bd[pb] * sphi + b[pb] * sphid) + S.pqc[2 * r + 1] * (bd[qb] * spsi + b[qb] * spsid);
#pragma unroll
        for (int w = 0; w < SB_WDL(r); w++) { const int e = SB_EOL(r) + w; acc += S.ec[e] * bd[SB_E_B(S.ei[e])] * U[SB_E_COL(S.ei[e])]; }
        const int i = SB_ROW(r);
        if (i < SB_N) out[i] = acc;
    }
    if (!SPL) __syncwarp();
}


#if defined(SB_WARP_PROF) && !defined(SB_SPLIT_PROF)
#define SB_SPLIT_PROF 1
#endif
#ifdef SB_SPLIT_PROF
// cycle accounting of one attempt by phase (debug builds only: scripts/split_prof.py); slot i = cycles between mark i-1 and mark i, thread 0 of each CTA
__device__ unsigned long long sb_split_prof[24];
#define SB_PROF_DECL long long pc_ = clock64(); unsigned long long pf_[20] = {0};
#define SB_PROF(i) { const long long c_ = clock64(); pf_[i] += (unsigned long long)(c_ - pc_); pc_ = c_; }
#define SB_PROF_PARM , unsigned long long* pf_ = nullptr, long long* pcp_ = nullptr
#define SB_PROF_PASS , pf_, &pc_
#define SB_PROF_F(i) { const long long c_ = clock64(); pf_[i] += (unsigned long long)(c_ - *pcp_); *pcp_ = c_; }
#ifdef SB_WARP_PROF
#define SB_WPROF(i) if (!BATCH && !GROUP) SB_PROF(i)
#define SB_WPROF_PASS , ((!BATCH && !GROUP) ? pf_ : nullptr), &pc_
#else
#define SB_WPROF(i)
#define SB_WPROF_PASS
#endif
#define SB_PROF_OUT if (tid == 0) { for (int i_ = 0; i_ < 16; i_++) { atomicAdd(&sb_split_prof[i_], pf_[i_]); pf_[i_] = 0; } atomicAdd(&sb_split_prof[20], (unsigned long long)(naccept + nreject)); } \
    if (tid == 1) { for (int i_ = 16; i_ < 20; i_++) { atomicAdd(&sb_split_prof[i_], pf_[i_]); pf_[i_] = 0; } } /* (slots 16..19: marks taken on thread 1) */
#else
#define SB_PROF_DECL
#define SB_WPROF(i)
#define SB_WPROF_PASS
#define SB_PROF_PARM
#define SB_PROF_PASS
#define SB_PROF_F(i)
#define SB_PROF(i)
#define SB_PROF_OUT
#endif
// Factor B = x·I − J_local(b): phase-0 elimination paths (zero fill, no pivoting: the chains have sign-skew off-diagonals
// and non-negative damping, so pivots only grow), then the top = root paths + explicit pivoted inverse of the top blocks.
// Leaves 1/pivot in di, multipliers in mm, (entry towards the parent)/pivot in up, block inverses in blk.
template <bool SPL = false>
__device__ __forceinline__ void sb_factor(const SbLane& S, double x, const double* b, double* di, double* up, double* mm, double* blk, int lane, int rb = 0, bool elim = true SB_PROF_PARM) {
    if (!SPL) { // (split kernel: every multiplier slot is either rewritten by the scatter below or structurally zero, so mm is cleared once per mode, and blk by
                // warp 0 after the attempt's last solve -- see sb_integrate_split_kernel)
#pragma unroll
        for (int r = 0; r < SB_NRL; r++) { const int i = SB_ROW(r); if (i < SB_N) mm[i] = 0; }
        for (int i = lane; i < SB_TOPSTORE; i += SB_WARP) blk[i] = 0;
        __syncwarp(); // the scatter below writes multipliers of rows other lanes have just cleared
    }
#pragma unroll
    for (int r = 0; r < SB_NRL; r++) {
        const int i = SB_ROW(r);
        double dacc = x, uacc = 0;
        // branch-free: all loads first, then selects and a predicated store per entry (adding 0.0 leaves the sums bit for bit as they were;
        // a branch per entry kept each load behind the previous entry's reconvergence point)
        double v[SB_WMAX]; // (round 0 is the widest)
#pragma unroll
        for (int w = 0; w < SB_WDL(r); w++) { const int e = SB_EOL(r) + w; v[w] = -S.ec[e] * b[SB_E_B(S.ei[e])]; }
#pragma unroll
        for (int w = 0; w < SB_WDL(r); w++) {
            const int e = SB_EOL(r) + w;
            const unsigned ix = S.ei[e];
            const int kind = SB_E_KIND(ix), tgt = SB_E_TGT(ix);
            dacc += (kind == 0) ? v[w] : 0.0;
            uacc += (kind == 1) ? v[w] : 0.0;
            if (kind >= 2 && S.ec[e] != 0.0) { // multiplier (target < N) or top-block entry (target >= N: blk follows mm)
#if SB_UNIQUE_TARGETS
                mm[tgt] = v[w];
#else
                mm[tgt] += v[w];
#endif
            }
        }
        if (i < SB_N) { di[i] = dacc; up[i] = uacc; } // di holds the raw diagonal until the vertex is eliminated
    }
    if (SPL) sb_rows_sync(); else __syncwarp();
    if (SPL) { SB_PROF_F(12) }
    if (SPL && !elim) return; // split kernel: the eliminations below are one warp's work (the paths and the top block live on lanes, not rows)
    // phase 0: every vertex but the last of a path is finished here (multiplier and parent coupling scaled by 1/pivot); the
    // last one keeps its raw multiplier for the owner of its parent, which forms the Schur term mm_raw * (up/pivot)
#pragma unroll
    for (int rd = 0; rd < SB_PR; rd++) {
        const int start = SB_P_START(S.ph[rd]), len = SB_P_LEN(S.ph[rd]);
        if (len > 0) {
            double* dip = di + start; double* mmp = mm + start; double* upp = up + start;
#if SB_PATH_CONT
            // Pivots of the tridiagonal path as ratios of continuants: p_i = a_i p_{i-1} − (m_{i-1} u_{i-1}) p_{i-2}, d_i = p_i / p_{i-1}.  The dependent
            // chain is ONE fma per vertex (the product with p_{i-2} is ready a step earlier) instead of multiply, fma and a Newton reciprocal; the
            // reciprocals of the p_i are independent of one another.  Sign-skew off-diagonals (m u < 0) and non-negative damping make every term of the
            // recurrence positive: no cancellation.  Growth is bounded by (1/(γ dt) + damping)^SB_PL, far inside the double range for SB_PL <= 12.
            double pp[SB_PL], mr[SB_PL], ur[SB_PL];
#pragma unroll
            for (int pos = 0; pos < SB_PL; pos++) { mr[pos] = 0; ur[pos] = 0; pp[pos] = 1.0; if (pos < len) { mr[pos] = mmp[pos]; ur[pos] = upp[pos]; pp[pos] = dip[pos]; } }
            {
                double pm2 = 1.0, pm1 = pp[0];
#pragma unroll
                for (int pos = 1; pos < SB_PL; pos++) {
                    const double p = fma(pp[pos], pm1, -((mr[pos - 1] * ur[pos - 1]) * pm2)); // (past the end of the path: 1·p_{i-1} − 0)
                    pp[pos] = p; pm2 = pm1; pm1 = p;
                }
            }
            double dv[SB_PL]; // 1/d_pos = p_{pos-1} / p_pos: straight-line, so that the reciprocals pipeline (past the end of the path: unused)
            dv[0] = sb_rcp(pp[0]);
#pragma unroll
            for (int pos = 1; pos < SB_PL; pos++) dv[pos] = pp[pos - 1] * sb_rcp(pp[pos]);
#pragma unroll
            for (int pos = 0; pos < SB_PL; pos++) if (pos < len) dip[pos] = dv[pos];
#pragma unroll
            for (int pos = 0; pos < SB_PL; pos++) { const double uv = ur[pos] * dv[pos]; if (pos < len) upp[pos] = uv; }
#pragma unroll
            for (int pos = 0; pos < SB_PL; pos++) { const double mv = mr[pos] * dv[pos]; if (pos + 1 < len) mmp[pos] = mv; } // (the last vertex keeps its raw multiplier for the owner of its parent)
#else
            double dinv = sb_rcp(dip[0]);
            dip[0] = dinv;
#pragma unroll
            for (int pos = 1; pos < SB_PL; pos++) {
                if (pos < len) {
                    const double mr = mmp[pos - 1], u = upp[pos - 1] * dinv;
                    mmp[pos - 1] = mr * dinv;
                    upp[pos - 1] = u;
                    dinv = sb_rcp(fma(-mr, u, dip[pos]));
                    dip[pos] = dinv;
                }
            }
            upp[len - 1] *= dinv;
#endif
        }
    }
    __syncwarp();
    if (SPL) { SB_PROF_F(13) }
    // top: root paths (fed by phase-0 paths only) ...
#pragma unroll
    for (int rd = 0; rd < SB_TR; rd++) { // (no such paths in the ΛCDM / w0waCDM models: SB_TR = 0)
        const int start = SB_P_START(S.rh[rd]), len = SB_P_LEN(S.rh[rd]);
        if (len > 0) {
            double* dip = di + start; double* mmp = mm + start; double* upp = up + start;
            double dj = dip[0];
#pragma unroll
            for (int c = 0; c < 3; c++) { const int ch = SB_KID(S.rk[rd], c); if (ch != SB_NOKID) { const double mr = mm[ch]; mm[ch] = mr * di[ch]; dj = fma(-mr, up[ch], dj); } }
            double dinv = sb_rcp(dj);
            dip[0] = dinv;
#pragma unroll
            for (int pos = 1; pos < SB_TL; pos++) {
                if (pos < len) {
                    const double mr = mmp[pos - 1], u = upp[pos - 1] * dinv;
                    mmp[pos - 1] = mr * dinv;
                    upp[pos - 1] = u;
                    dinv = sb_rcp(fma(-mr, u, dip[pos]));
                    dip[pos] = dinv;
                }
            }
            upp[len - 1] = 0;
        }
    }
    // ... and top vertices (block members and single roots): gather the children's Schur terms into the diagonal; singles
    // are done with that; blocks: Gauss-Jordan with partial pivoting, one matrix row [A | I] per lane (lane = block*8 + row),
    // rows exchanged by warp shuffles within the 8-lane group -> explicit inverse (block solves become mat-vecs)
    {
        const int bb = lane >> 3, i = lane & 7;
        int nb = 0, off = 0;
        if (bb < SB_NTOP) { nb = sb_top_n[bb]; off = sb_top_off[bb]; }
        double dj = 1.0;
        if (S.tv != SB_NOKID) {
            dj = di[S.tv];
            // branch-free gather: an absent child reads this lane's own vertex instead and its term is dropped by a select (a branch per child
            // kept each child's loads behind the previous child's reconvergence point)
            double mrv[3], div[3], upv[3];
#pragma unroll
            for (int c = 0; c < 3; c++) { const int ch = SB_KID(S.dk, c); const int cs = (ch != SB_NOKID) ? ch : (int)S.tv; mrv[c] = mm[cs]; div[c] = di[cs]; upv[c] = up[cs]; }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const int ch = SB_KID(S.dk, c);
                const double tnew = fma(-mrv[c], upv[c], dj);
                if (ch != SB_NOKID) mm[ch] = mrv[c] * div[c];
                dj = (ch != SB_NOKID) ? tnew : dj;
            }
            if (i >= nb) { di[S.tv] = sb_rcp(dj); up[S.tv] = 0; }
        }
        if (SPL) { SB_PROF_F(14) }
#if SB_GJ_INPLACE
        // In-place Gauss-Jordan: the column of the pivot is overwritten by the corresponding column of the inverse, so only the
        // SB_TOPMAX entries of the pivot row travel per step (the augmented form [A | I] moves twice as many).  Rows are not
        // exchanged: lane `who` simply acts as row kx, which leaves the inverse with its columns in pivot order -- column kx of the
        // register row is column p_kx (= who of step kx) of the inverse; the scatter at the end undoes that.
        double Ar[SB_TOPMAX];
#pragma unroll
        for (int j = 0; j < SB_TOPMAX; j++) { // (unconditional loads from a clamped index, then selects)
            const bool inb = i < nb && j < nb;
            const double bv = blk[inb ? off + i * nb + j : 0];
            Ar[j] = (i == j) ? ((i < nb) ? dj : 1.0) : (inb ? bv : 0.0);
        }
        int myrow = -1; // pivot column this lane's row was used for (= its row index in the inverse)
        unsigned perm = 0; // who of every step, 3 bits each
        const int base = lane & ~7;
#pragma unroll
        for (int kx = 0; kx < SB_TOPMAX; kx++) {
#if SB_GJ_REDUX && SB_NTOP == 1
            // one top block (lanes 0..4): a single full-warp REDUX on the high word of |a| (exponent + 16 mantissa bits; the lower
            // row wins a tie; candidates that agree to 1.5e-5 relative are equally good pivots).  (A REDUX per 8-lane group, which
            // several blocks would need, measured slower than the shuffle search below.)
            unsigned key = 0;
            if (myrow < 0 && i < SB_TOPMAX && bb == 0) key = (((unsigned)__double2hiint(Ar[kx]) & 0x7FFFFFF0u)) | (unsigned)(8 - i);
            const double cinv = sb_rcp(Ar[kx]); // every candidate's reciprocal, formed while the search is in flight (same value as rcp of the broadcast pivot)
            const int who = 8 - (int)(__reduce_max_sync(SB_FULL, key) & 15u);
#else
            const double cinv = sb_rcp(Ar[kx]);
            double best = (myrow < 0 && i < SB_TOPMAX) ? fabs(Ar[kx]) : -1.0; int who = i;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) { const double ob = __shfl_xor_sync(SB_FULL, best, o); const int ow = __shfl_xor_sync(SB_FULL, who, o); if (ob > best || (ob == best && ow < who)) { best = ob; who = ow; } }
#endif
            const bool isp = (i == who);
            if (isp) myrow = kx;
            perm |= (unsigned)who << (3 * kx);
            const double inv = __shfl_sync(SB_FULL, cinv, base + who);
            const double l = isp ? 0.0 : Ar[kx];
#pragma unroll
            for (int j = 0; j < SB_TOPMAX; j++) {
                if (j == kx) continue;
                const double pa = __shfl_sync(SB_FULL, Ar[j], base + who) * inv;
                if (isp) Ar[j] = pa; else Ar[j] -= l * pa;
            }
            Ar[kx] = isp ? inv : -l * inv;
        }
        if (bb < SB_NTOP && myrow >= 0 && myrow < nb) {
#pragma unroll
            for (int j = 0; j < SB_TOPMAX; j++) { const int pj = (perm >> (3 * j)) & 7; if (pj < nb) blk[off + myrow * nb + pj] = Ar[j]; }
        }
    }
#else
        double Ar[SB_TOPMAX], Ir[SB_TOPMAX];
#pragma unroll
        for (int j = 0; j < SB_TOPMAX; j++) {
            Ar[j] = (i == j) ? ((i < nb) ? dj : 1.0) : ((i < nb && j < nb) ? blk[off + i * nb + j] : 0.0);
            Ir[j] = (i == j) ? 1.0 : 0.0;
        }
        int myrow = -1; // pivot column this lane's row was used for (= its row index in the inverse)
        const int base = lane & ~7;
#pragma unroll
        for (int kx = 0; kx < SB_TOPMAX; kx++) {
            // pivot search among rows not used yet
            double best = (myrow < 0 && i < SB_TOPMAX) ? fabs(Ar[kx]) : -1.0; int who = i;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) { const double ob = __shfl_xor_sync(SB_FULL, best, o); const int ow = __shfl_xor_sync(SB_FULL, who, o); if (ob > best || (ob == best && ow < who)) { best = ob; who = ow; } }
            if (i == who) myrow = kx;
            const double piv = __shfl_sync(SB_FULL, Ar[kx], base + who);
            const double inv = sb_rcp(piv);
            const double l = (i == who) ? 0.0 : Ar[kx];
#pragma unroll
            for (int j = 0; j < SB_TOPMAX; j++) {
                const double pa = __shfl_sync(SB_FULL, Ar[j], base + who) * inv, pi_ = __shfl_sync(SB_FULL, Ir[j], base + who) * inv;
                if (i == who) { Ar[j] = pa; Ir[j] = pi_; } else { Ar[j] -= l * pa; Ir[j] -= l * pi_; }
            }
        }
        if (bb < SB_NTOP && myrow >= 0 && myrow < nb) {
#pragma unroll
            for (int j = 0; j < SB_TOPMAX; j++) if (j < nb) blk[off + myrow * nb + j] = Ir[j];
        }
    }
#endif
    __syncwarp();
    if (SPL) { SB_PROF_F(15) }
}

// r <- B^{-1} r for NR right-hand sides at once (independent recurrences interleave: the solve is latency-bound), in three
// warp-synchronous steps: forward along the phase-0 paths (registers) | top: root paths forward + backward, top blocks =
// gather the children's terms, exchange by shuffles, mat-vec with the explicit inverse | backward along the phase-0 paths.
template <int NR>
__device__ __forceinline__ void sb_bsolve(const SbLane& S, double* const (&rr)[NR], const double* __restrict__ di, const double* __restrict__ up, const double* __restrict__ mm, const double* __restrict__ blk, int lane) {
    // KEEP: the forward-substitution results of the phase-0 paths stay in registers across the top step (the shared-memory data
    // path is the busiest unit of this kernel: this saves the store and the reload of every interior path vertex); only the last
    // vertex of a path, which the owner of its parent gathers, is stored.  Positions are counted from the END of the path
    // (q = 0 is the last vertex) in both sweeps, so the register arrays of the two sweeps line up for paths of any length.
    constexpr bool KEEP = (SB_PR * SB_PL * NR <= SB_KEEP_MAX);
    double fw[KEEP ? SB_PR : 1][KEEP ? NR : 1][KEEP ? SB_PL : 1];
#pragma unroll
    for (int rd = 0; rd < SB_PR; rd++) {
        // The recurrence runs unpredicated over the longest path length with predicated loads: past the end of a shorter path
        // (or on a lane without a path) it computes zeros that are never stored.  This keeps the dependent DFMA chain free of
        // predicated register moves and lets all loads issue from one base register.
        const int start = SB_P_START(S.ph[rd]), len = SB_P_LEN(S.ph[rd]);
        if (KEEP) {
            const int end = start + len - 1;
            const double* mme = mm + end;
#pragma unroll
            for (int a = 0; a < NR; a++) {
                double* re = rr[a] + end;
                double acc = 0.0;
#pragma unroll
                for (int q = SB_PL - 1; q >= 0; q--) {
                    // Unconditional loads: positions before the first vertex of a shorter path (and lanes without a path) read
                    // neighbouring shared-memory values (always inside the CTA's allocation: the arrays swept here start at least 2N
                    // doubles into it) and compute garbage that is never stored; the accumulator is cleared at the first vertex,
                    // which has no predecessor, so the garbage does not reach the path.  (A predicated load costs four
                    // instructions -- load, clear, two moves -- which made this sweep a fifth of the kernel's instruction count.)
                    const double m_ = mme[-q - 1], r_ = re[-q];
                    if (q + 1 >= len) acc = 0.0;
                    acc = fma(-m_, acc, r_);
                    fw[rd][a][q] = acc;
                }
                if (len > 1) re[0] = fw[rd][a][0];
            }
        } else {
            const double* mmp = mm + start;
#pragma unroll
            for (int a = 0; a < NR; a++) {
                double* rp = rr[a] + start;
                double av[SB_PL];
                double acc = rp[0];
#pragma unroll
                for (int pos = 1; pos < SB_PL; pos++) { acc = fma(-mmp[pos - 1], acc, rp[pos]); av[pos] = acc; } // past the end of the path: garbage, never stored
#pragma unroll
                for (int pos = 1; pos < SB_PL; pos++) if (pos < len) rp[pos] = av[pos];
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int rd = 0; rd < SB_TR; rd++) { // (no such paths in the ΛCDM / w0waCDM models: SB_TR = 0)
        const int start = SB_P_START(S.rh[rd]), len = SB_P_LEN(S.rh[rd]);
        if (len > 0) {
            const double *mmp = mm + start, *upp = up + start, *dip = di + start;
#pragma unroll
            for (int a = 0; a < NR; a++) {
                double* rp = rr[a] + start;
                double av[SB_TL];
                double acc = rp[0];
#pragma unroll
                for (int c = 0; c < 3; c++) { const int ch = SB_KID(S.rk[rd], c); if (ch != SB_NOKID) acc = fma(-mm[ch], rr[a][ch], acc); }
                av[0] = acc;
#pragma unroll
                for (int pos = 1; pos < SB_TL; pos++) if (pos < len) { acc = fma(-mmp[pos - 1], acc, rp[pos]); av[pos] = acc; }
                double xv = 0.0;
#pragma unroll
                for (int pos = SB_TL - 1; pos >= 0; pos--) if (pos < len) { xv = fma(-upp[pos], xv, av[pos] * dip[pos]); av[pos] = xv; }
#pragma unroll
                for (int pos = 0; pos < SB_TL; pos++) if (pos < len) rp[pos] = av[pos];
            }
        }
    }
#if SB_TOPFAST && SB_NTOP == 1
    { // top vertices, one top block of SB_TOPMAX rows on lanes 0..SB_TOPMAX-1 and single roots on the lanes after the block group.
      // No predicated loads (four instructions each): absent children point at a top-block vertex, whose multiplier mm is always 0;
      // the block row and the single-root pivot are loaded by every top lane and the unused one is dropped by a select.
        const int i = lane & 7, base = lane & ~7;
        const bool mine = S.tv != SB_NOKID, inblk = lane < SB_TOPMAX;
        double g[NR], brow[SB_TOPMAX];
#pragma unroll
        for (int a = 0; a < NR; a++) g[a] = 0;
        double dtv = 0;
        if (mine) {
            dtv = di[S.tv];
#pragma unroll
            for (int j = 0; j < SB_TOPMAX; j++) brow[j] = blk[i * SB_TOPMAX + j]; // single roots: values of the neighbouring array, dropped below
#pragma unroll
            for (int a = 0; a < NR; a++) {
                double ga = rr[a][S.tv];
#pragma unroll
                for (int c = 0; c < 3; c++) { const int ch = SB_KID(S.dkz, c); ga = fma(-mm[ch], rr[a][ch], ga); }
                g[a] = ga;
            }
        }
#pragma unroll
        for (int a = 0; a < NR; a++) {
            double xm = 0;
#pragma unroll
            for (int j = 0; j < SB_TOPMAX; j++) xm = fma(brow[j], __shfl_sync(SB_FULL, g[a], base + j), xm);
            const double xi = inblk ? xm : g[a] * dtv; // single root: forward and backward substitution are one division by the pivot
            if (mine) rr[a][S.tv] = xi;
        }
    }
#else
    { // top vertices: block members (lane = block*8 + row) and single roots gather their children's terms in one pass
        const int bb = lane >> 3, i = lane & 7, base = lane & ~7;
        int nb = 0, off = 0;
        if (bb < SB_NTOP) { nb = sb_top_n[bb]; off = sb_top_off[bb]; }
        const bool mine = S.tv != SB_NOKID, inblk = i < nb;
        double g[NR], brow[SB_TOPMAX];
        const double dtv = (mine && !inblk) ? di[S.tv] : 0.0;
#pragma unroll
        for (int j = 0; j < SB_TOPMAX; j++) brow[j] = (inblk && j < nb) ? blk[off + i * nb + j] : 0.0;
#pragma unroll
        for (int a = 0; a < NR; a++) {
            g[a] = 0;
            if (mine) {
                g[a] = rr[a][S.tv];
#pragma unroll
                for (int c = 0; c < 3; c++) { const int ch = SB_KID(S.dk, c); if (ch != SB_NOKID) g[a] = fma(-mm[ch], rr[a][ch], g[a]); }
            }
        }
#pragma unroll
        for (int a = 0; a < NR; a++) {
            double xi = g[a] * dtv; // single root: forward and backward substitution are one division by the pivot
            if (SB_NTOP > 0) {
#pragma unroll
                for (int j = 0; j < SB_TOPMAX; j++) xi = fma(brow[j], __shfl_sync(SB_FULL, g[a], base + j), xi);
            }
            if (mine) rr[a][S.tv] = xi;
        }
    }
#endif
    __syncwarp();
#pragma unroll
    for (int rd = 0; rd < SB_PR; rd++) {
        // backward along the path from its last vertex (q = 0) to its first, again unpredicated: offsets count down from the end
        const int start = SB_P_START(S.ph[rd]), len = SB_P_LEN(S.ph[rd]), par = SB_P_PAR(S.ph[rd]);
        const int end = start + len - 1;
        const double *upe = up + end, *die = di + end;
#pragma unroll
        for (int a = 0; a < NR; a++) {
            double* re = rr[a] + end;
            double xs[SB_PL];
            double xv = (par == SB_NOPAR) ? 0.0 : rr[a][par];
#pragma unroll
            for (int q = 0; q < SB_PL; q++) { const double r_ = KEEP ? fw[KEEP ? rd : 0][KEEP ? a : 0][KEEP ? q : 0] : re[-q]; xv = fma(-upe[-q], xv, r_ * die[-q]); xs[q] = xv; } // q >= len: garbage, never stored
#pragma unroll
            for (int q = 0; q < SB_PL; q++) if (q < len) re[-q] = xs[q];
        }
    }
    __syncwarp();
}

__device__ __forceinline__ double sb_wcorr(double k, double zp, double c1, double zq, double c2) { return k + (zp * c1 + zq * c2); } // k + Z c (last stage)
__device__ __forceinline__ void sb_hub_dots(const SbLane& S, const double* b, const double* r, double& s1, double& s2, int lane) {
    double a = 0, c = 0;
#pragma unroll
    for (int t = 0; t < SB_TPH; t++) a += S.phc[t] * b[SB_HI16(S.phi[t])] * r[SB_LO16(S.phi[t])];
#pragma unroll
    for (int t = 0; t < SB_TPS; t++) c += S.psc[t] * b[SB_HI16(S.psi[t])] * r[SB_LO16(S.psi[t])];
    s1 = warp_sum(a); s2 = warp_sum(c);
}

// ------------------------------------------------------------------------------------------------ CMB source functions
// ST, SE (reference src/models/cosmologies.jl:99-104) and the lensing source Sψ (:105) from ONE saved state u (natural order).
// Derivatives of unknowns are expanded through the ODE itself (u̇ = J u, ü = J̇ u + J u̇), as MTK does symbolically.  The pieces
// below are shared by the fused evaluation inside the integrator (sb_source_point: the reference evaluates S in solvept's
// output_func, src/observables/fourier.jl:272-278) and by the stand-alone kernel over saved states (sb_source_kernel), so that
// both produce the same bits from the same state.
// Per-τ background row srcbg[it][0..5] = κ̇, κ̈, κ⃛, exp(−κ), τ0 − τ, spare;  then β_m (NBETA) and flow-derivative dβ_m/dτ (NBETA).
#define SB_SRCBG_STRIDE (8 + 2 * SB_NBETA)
__device__ __forceinline__ double sb_row_dot(int i, const double* b, const double* u) { // (J_local u)_i
    double acc = 0;
    for (int e = sb_rowptr[i]; e < sb_rowptr[i + 1]; e++) acc += sb_coef[e] * b[sb_bidx[e]] * u[sb_col[e]];
    return acc;
}
__device__ __forceinline__ double sb_row_ddot(int i, const double* b, const double* bd, const double* u, const double* ud) { // (J̇_local u + J_local u̇)_i
    double acc = 0;
    for (int e = sb_rowptr[i]; e < sb_rowptr[i + 1]; e++) acc += sb_coef[e] * (bd[sb_bidx[e]] * u[sb_col[e]] + b[sb_bidx[e]] * ud[sb_col[e]]);
    return acc;
}
__device__ __forceinline__ double sb_hub_seq(int v, const double* b, const double* u) { // hub functional v (2: Φ̇ = φᵀu, 3: Ψ = ψᵀu)
    double acc = 0;
    for (int t = sb_hptr[v]; t < sb_hptr[v + 1]; t++) acc += sb_hcoef[t] * b[sb_hb[t]] * u[sb_hidx[t]];
    return acc;
}
// the scalar tail, run by ONE thread: hub columns into u̇, Ψ̇, Π and its derivatives, the sources.  ud holds J_local u on entry.
__device__ __forceinline__ void sb_source_tail(const double* sb, double k, double tau, double taurec, int scale_k, int nS, const double* b, const double* bd, const double* u, double* ud,
                                               double Phd, double Psi, double* out, int ostride) {
    for (int t = sb_hptr[0]; t < sb_hptr[1]; t++) ud[sb_hidx[t]] += sb_hcoef[t] * b[sb_hb[t]] * Phd;
    for (int t = sb_hptr[1]; t < sb_hptr[2]; t++) ud[sb_hidx[t]] += sb_hcoef[t] * b[sb_hb[t]] * Psi;
    double Psid = 0;
    for (int t = sb_hptr[3]; t < sb_hptr[4]; t++) Psid += sb_hcoef[t] * (bd[sb_hb[t]] * u[sb_hidx[t]] + b[sb_hb[t]] * ud[sb_hidx[t]]);
    // second derivatives of F2, G0, G2 (rows without hub terms): ü_i = J_i(ḃ) u + J_i(b) u̇
    const int rows3[3] = {SB_I_F2, SB_I_G0, SB_I_G2};
    double Pig = 0, Pigd = 0, Pigdd = 0;
    for (int r = 0; r < 3; r++) {
        const int i = rows3[r];
        const double acc = sb_row_ddot(i, b, bd, u, ud);
        Pig += u[i]; Pigd += ud[i]; Pigdd += acc;
    }
    const double kd = sb[0], kdd = sb[1], kddd = sb[2], ek = sb[3], chi = sb[4];
    const double v = -kd * ek, vd = (-kdd + kd * kd) * ek, vdd = (-kddd + 3 * kd * kdd - kd * kd * kd) * ek; // v = d/dτ e^{−κ}
    const double thb = u[SB_I_TB], thbd = ud[SB_I_TB];
    double ST = v * (u[SB_I_F0] / 4 + Psi + Pig / 16) + ek * (Psid + Phd) + (vd * thb + v * thbd) / (k * k) + 3.0 / (16 * k * k) * (vdd * Pig + 2 * vd * Pigd + v * Pigdd);
    double SE = 3.0 / 16.0 * v * Pig / ((k * chi) * (k * chi));
    if (scale_k) { ST *= k; SE *= k * k; } // the LOS integrator is fed k·ST and k²·SE (src/observables/angular.jl:293)
    out[0] = ST;
    out[ostride] = SE;
    if (nS > 2) out[2 * ostride] = (tau >= taurec) ? -(Psi + u[SB_I_PHI]) * (tau - taurec) / (chi + tau - taurec) / chi : 0.0 * Psi; // (0·Ψ: a failed mode stays NaN) // τ0 = χ + τ
}
// Coalesced store of the window of source values: lane 0 has put the sources of save index `it` into sbuf[s][it % SB_SWIN]; when the
// window is full (or the mode's last save time is reached) the lanes write it as contiguous runs of S[s][w0 .. w0+cnt).  wstart: first
// valid slot of the window (non-zero only right after a parked mode was resumed in the middle of a window).
__device__ __forceinline__ void sb_source_flush(const double* sbuf, double* Sout, int nS, int nsave, int it, int& wstart, bool force, int lane) {
    __syncwarp();
    if ((it & (SB_SWIN - 1)) == SB_SWIN - 1 || it == nsave - 1 || force) {
        const int w0 = it & ~(SB_SWIN - 1), cnt = (it & (SB_SWIN - 1)) + 1;
        for (int q = lane; q < nS * SB_SWIN; q += SB_WARP) {
            const int s = q / SB_SWIN, j = q - s * SB_SWIN;
            if (j >= wstart && j < cnt) Sout[(size_t)s * nsave + w0 + j] = sbuf[q];
        }
        wstart = 0;
        __syncwarp();
    }
}

// Pieces shared verbatim by the warp-per-mode and the split kernel.  Sums of two products are written with explicit fma so that the two
// kernels contract them the same way (bit-identical results are tested: tests/test_gpu_parity.py::test_split_kernel_is_bit_identical).
__device__ __forceinline__ double sb_dd_term(double c, double bd, double u, double b, double ud, double acc) { return fma(c, fma(bd, u, b * ud), acc); } // acc + c·(ḃu + bu̇)
__device__ __forceinline__ void sb_source_close(const double* __restrict__ sb, double k, double tau, double taurec, int scale_k, int nS, double Phd, double Psi, double Psid, double Pig, double Pigd,
                                                double Pigdd, double thb, double thbd, double F0, double Phi, double* out, int ostride) {
    const double kd = __ldg(sb), kdd = __ldg(sb + 1), kddd = __ldg(sb + 2), ek = __ldg(sb + 3), chi = __ldg(sb + 4);
    const double v = -kd * ek, vd = fma(kd, kd, -kdd) * ek, vdd = (fma(3 * kd, kdd, -kddd) - kd * kd * kd) * ek; // v = d/dτ e^{−κ}
    double ST = v * (F0 / 4 + Psi + Pig / 16);
    ST = fma(ek, Psid + Phd, ST);
    ST += fma(vd, thb, v * thbd) / (k * k);
    ST = fma(3.0 / (16 * k * k), fma(vdd, Pig, fma(2 * vd, Pigd, v * Pigdd)), ST);
    double SE = 3.0 / 16.0 * v * Pig / ((k * chi) * (k * chi));
    if (scale_k) { ST *= k; SE *= k * k; } // the LOS integrator is fed k·ST and k²·SE (src/observables/angular.jl:293)
    out[0] = ST;
    out[ostride] = SE;
    if (nS > 2) out[2 * ostride] = (tau >= taurec) ? -(Psi + Phi) * (tau - taurec) / (chi + tau - taurec) / chi : 0.0 * Psi; // (0·Ψ: a failed mode stays NaN); τ0 = χ + τ
}
// Rodas5P dense output (4th order) at θ = (ts − t)/dt from u_n, u_{n+1} and the three combinations a_m = Σ_j H_mj k_j
__device__ __forceinline__ double sb_dense(double th, double t1, double u, double un, double a1, double a2, double a3) {
    return fma(th, fma(t1, fma(th, fma(th, a3, a2), a1), un), t1 * u);
}

// Warp-cooperative evaluation at one save time inside the integrator, on the integrator's own register-resident schedule (the
// generated CSR tables of the stand-alone kernel live in global memory: walking them from here costs an L2 round trip per entry,
// which made a source point as expensive as half a Rosenbrock attempt).  u: the saved state in the integrator's order, ud/b/bd:
// scratch (all in the warp's shared memory); out[s * ostride] receives source s (written by lane 0).  Same algebra as
// sb_source_tail; the sums run in a different order (lanes + butterfly), so results agree with it to rounding, not bit for bit.
__device__ __forceinline__ void sb_source_point(const SbLane& S, const double* __restrict__ sb, const double* kp, double k, double tau, double taurec, int scale_k, int nS, const double* u,
                                                double* ud, double* b, double* bd, double* out, int ostride, int lane) {
#pragma unroll
    for (int r = 0; r < SB_NBR; r++) {
        const int m = r * 32 + lane;
        if (m < SB_NB) {
            const int be = SB_LO16(S.bp[r]);
            const double kk = kp[SB_HI16(S.bp[r])];
            b[m] = kk * __ldg(sb + 8 + be);
            bd[m] = kk * __ldg(sb + 8 + SB_NBETA + be);
        }
    }
    __syncwarp();
    double hub[2];
    sb_eval_f<false>(S, b, u, ud, lane, nullptr, 0.0, nullptr, hub); // u̇ = J u (hub columns included), Φ̇, Ψ
    const double Phd = hub[0], Psi = hub[1];
    // Ψ̇ = ψ̇ᵀu + ψᵀu̇ and Π̈ = Σ_{F2,G0,G2} (J̇_i u + J_i u̇) (rows without hub terms), reduced together
    double a = 0, pdd = 0;
#pragma unroll
    for (int t = 0; t < SB_TPS; t++) a = sb_dd_term(S.psc[t], bd[SB_HI16(S.psi[t])], u[SB_LO16(S.psi[t])], b[SB_HI16(S.psi[t])], ud[SB_LO16(S.psi[t])], a);
#pragma unroll
    for (int r = 0; r < SB_R; r++) {
        const int i = r * 32 + lane;
        if (i == SB_J_F2 || i == SB_J_G0 || i == SB_J_G2) {
#pragma unroll
            for (int w = 0; w < SB_WDR(r); w++) { const int e = SB_EOFF(r) + w; pdd = sb_dd_term(S.ec[e], bd[SB_E_B(S.ei[e])], u[SB_E_COL(S.ei[e])], b[SB_E_B(S.ei[e])], ud[SB_E_COL(S.ei[e])], pdd); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(SB_FULL, a, o); pdd += __shfl_xor_sync(SB_FULL, pdd, o); }
    if (lane == 0)
        sb_source_close(sb, k, tau, taurec, scale_k, nS, Phd, Psi, a, u[SB_J_F2] + u[SB_J_G0] + u[SB_J_G2], ud[SB_J_F2] + ud[SB_J_G0] + ud[SB_J_G2], pdd, u[SB_J_TB], ud[SB_J_TB], u[SB_J_F0],
                        u[SB_J_PHI], out, ostride);
    __syncwarp();
}

// Basis at the stage times 2..6 of an attempt (slots 1..5; slot 0, the time t, is carried over from the previous attempt's last stage time).
// Deferred form of sb_basis_batch (SB_DEFER; warp-per-mode and split kernel): the rows are requested at the start of the attempt, the factorisation and the
// first solve need slot 0 only, and the wait + sweep come right before the first stage evaluation -- the L2 latency of the copies (≈1.5 k cycles, 6 % of a lone
// warp's attempt) is off the dependent chain.  One warp locates the slots and issues the table-row copies ...
__device__ __forceinline__ void sb_split_basis_issue(const SbTable& tb, double t, double dt, int jt, double* slotp, double* tabs, int* jend, int lane SB_PROF_PARM) {
    const unsigned mbar = sb_smem_u32(tabs + 6 * 4 * SB_NBETA);
    constexpr unsigned ROWB = 4 * SB_NBETA * 8;
#ifndef SB_EXP_NOFENCE
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the previous attempt's generic-proxy reads of the staging area come first
#endif
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(5 * ROWB) : "memory");
    if (lane >= 1 && lane < 6) {
        double* sp = slotp + lane * 8;
        const int j = sb_slot_locate(tb, fma(cc[lane], dt, t), jt, sp);
        const double* src = tb.tab + (size_t)__double_as_longlong(sp[0]) * 2 * SB_NBETA;
#ifdef SB_WARP_PROF
        if (pf_ && lane == 1) { SB_PROF_F(16) } // (debug: the search is done)
#endif
#ifndef SB_EXP_NOCOPY
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb_smem_u32(tabs + lane * 4 * SB_NBETA)), "l"(src), "r"(ROWB), "r"(mbar) : "memory");
#endif
        if (lane == 5) *jend = j; // interval of t + dt: the next attempt's starting hint
#ifdef SB_WARP_PROF
        if (pf_ && lane == 1) { SB_PROF_F(17) } // (debug: the copies are issued)
#endif
    }
    __syncwarp();
}
// ... and the sweeping warps (split kernel: all but warp 0, which runs the eliminations meanwhile; warp kernel: the warp itself) wait for the rows and
// sweep the slots, `widx` of `nw` each
__device__ __forceinline__ void sb_split_basis_sweep(const SbLane& S, const double* kp, double* bs, double* bd5, const double* slotp, const double* tabs, unsigned ph, int lane, int widx, int nw) {
    const unsigned mbar = sb_smem_u32(tabs + 6 * 4 * SB_NBETA);
    asm volatile("{\n .reg .pred p;\n SB_WAITS_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra SB_WAITS_%=;\n}" ::"r"(mbar), "r"(ph) : "memory");
    double kk[SB_NBR];
    int be[SB_NBR];
#pragma unroll
    for (int r = 0; r < SB_NBR; r++) { be[r] = SB_LO16(S.bp[r]); kk[r] = kp[SB_HI16(S.bp[r])]; }
#pragma unroll
    for (int s = 1; s < 6; s++) {
        if ((s - 1) % nw != widx) continue;
        const double* sp = slotp + s * 8;
        const double w1 = sp[1], w2 = sp[2], w3 = sp[3], w4 = sp[4];
        const double* n0 = tabs + s * 4 * SB_NBETA;
#pragma unroll
        for (int r = 0; r < SB_NBR; r++) {
            const int m = r * 32 + lane;
            if (m < SB_NB) {
                const double v0 = n0[be[r]], d0 = n0[SB_NBETA + be[r]], v1 = n0[2 * SB_NBETA + be[r]], d1 = n0[3 * SB_NBETA + be[r]];
                bs[s * SB_NB + m] = sb_hermite_val(kk[r], w1, v0, w2, d0, w3, v1, w4, d1);
                if (s == 5) bd5[m] = sb_hermite_der(kk[r], sp[5], v0, v1, sp[6], d0, sp[7], d1); // the next attempt's slot 0 derivative, if this one is accepted
            }
        }
    }
}
// Slot 0 (time t, value and derivative) from the table in global memory with the sweep's own expressions: for a mode that resumes a parked solve, whose
// slot 0 an uninterrupted solve would have carried over from the previous attempt's last stage time (same values bit for bit).
__device__ __forceinline__ int sb_basis_slot0(const SbLane& S, const SbTable& tb, double t, int jt, const double* kp, double* bs, double* bdv, double* slotp, int lane) {
    int j = jt;
    if (lane == 0) j = sb_slot_locate(tb, t, jt, slotp);
    j = __shfl_sync(SB_FULL, j, 0);
    __syncwarp();
    const double* sp = slotp;
    const double* n0 = tb.tab + (size_t)__double_as_longlong(sp[0]) * 2 * SB_NBETA;
#pragma unroll
    for (int r = 0; r < SB_NBR; r++) {
        const int m = r * 32 + lane;
        if (m < SB_NB) {
            const int be = SB_LO16(S.bp[r]);
            const double kk = kp[SB_HI16(S.bp[r])];
            const double v0 = __ldg(n0 + be), d0 = __ldg(n0 + SB_NBETA + be), v1 = __ldg(n0 + 2 * SB_NBETA + be), d1 = __ldg(n0 + 3 * SB_NBETA + be);
            bs[m] = sb_hermite_val(kk, sp[1], v0, sp[2], d0, sp[3], v1, sp[4], d1);
            bdv[m] = sb_hermite_der(kk, sp[5], v0, v1, sp[6], d0, sp[7], d1);
        }
    }
    __syncwarp();
    return j;
}
// Initial state of a mode (generated closed-form initial conditions on the background at τini).  One compiled copy for every integrator kernel: inlined,
// the several hundred generated expressions were contracted into FMAs differently in different kernels (nx = 8: two unknowns off by an ulp between the
// warp-per-mode and the split kernel), and results must not depend on which kernel runs a mode.
__device__ __noinline__ void sb_initial_state(const SbSpline& spl, const double* P, double t, double k, double* U) {
    double y[5];
    sb_spl_eval(spl, t, y, nullptr);
    sb_initial(t, k, y, P, U);
}
#ifdef SB_TIME_PROF
// wall-clock accounting of a scheduled launch (debug builds only: scripts/sched_prof.py): per list (start, end, time spent waiting for a parked mode), per mode (ns over its pieces)
__device__ unsigned long long sb_tp_list[4096 * 3];
__device__ unsigned long long sb_tp_mode[8192];
__device__ __forceinline__ unsigned long long sb_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
// Persistent kernel: one warp per k-mode (SB_WARPS_PER_CTA independent warps per CTA), modes pulled from an atomic work
// queue in the given order (host sorts by descending k, i.e. descending cost).  FP64 throughout.
// BATCH: every mode carries its own cosmology (A.cosmos[A.cosmo_of[mode]], staged in shared memory per mode) -- one launch over
// the (cosmology, mode) pairs of a parameter sweep; otherwise the single cosmology A.c0 is read from the kernel parameters.
// GROUP (sbm_solvept_lanes; BASELINE config 5, "dual-number lanes in the batched solve"): the CTA has A.G warps; warp j integrates
// the SAME mode for cosmology j (lane 0: the primal, lanes j >= 1: one parameter moved by δ_j) and all of them share one step
// controller: every attempt ends with one CTA barrier at which the lanes exchange (u_n, u_{n+1}, k_8); each warp then forms the same
// error norm over the primal AND the partials (u^j − u^0)/δ_j (OrdinaryDiffEq's norm of Dual numbers), so accept/reject
// and the next step are identical in all lanes.  With frozen, shared steps the lanes' results are smooth in the parameters and
// (u^j − u^0)/δ_j is the derivative of the discrete solution map -- what ForwardDiff computes -- up to O(δ).
template <bool BATCH, bool GROUP = false>
#ifdef SB_NOLB
__global__ void sb_integrate_kernel
#else
__global__ void __launch_bounds__(GROUP ? SB_WARP * SB_GROUP_MAX : SB_WARP * SB_WARPS_PER_CTA, GROUP ? 1 : (BATCH ? SB_MINBLOCKS_BATCH : SB_MINBLOCKS)) sb_integrate_kernel
#endif
    (const __grid_constant__ SbSolveArgs A) {
    extern __shared__ __align__(16) double sm_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
#if SB_WARPS_PER_CTA == 1
    double* const sm = GROUP ? sm_all + warp * SB_SM_DOUBLES : sm_all; // constant base: shared-memory accesses become [index.X8 + immediate], no address arithmetic
#else
    double* sm = sm_all + warp * SB_SM_DOUBLES;
#endif
    double* const xch = sm_all + (GROUP ? A.G : 0) * SB_SM_DOUBLES; // GROUP: header + [2 parities][G lanes][3][N] exchange of (u_n, u_{n+1}, k_8)
    double *u = sm + SB_SM_U, *U = sm + SB_SM_UU, *K = sm + SB_SM_K, *dT = sm + SB_SM_DT, *f0 = sm + SB_SM_F0, *di = sm + SB_SM_DI, *up = sm + SB_SM_UP,
           *mm = sm + SB_SM_MM, *Zp = sm + SB_SM_ZP, *Zq = sm + SB_SM_ZQ, *bs = sm + SB_SM_BS, *bdv = sm + SB_SM_BD, *blk = sm + SB_SM_BLK, *kp = sm + SB_SM_KP;
    double* const sbuf = sm + SB_SM_SBUF;
#if SB_TMA
    double* const tabs = sm + SB_SM_TAB;
    unsigned tpar = 0;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb_smem_u32(sm + SB_SM_MBAR)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
#endif
    const double reltol = A.reltol, abstol = A.abstol;
    const bool SRC = A.S != nullptr; // fused source evaluation at the save times (scratch: di = state in natural order, up = its derivative, bs = basis)
#ifdef SB_WARP_PROF
    SB_PROF_DECL // (debug builds: cycle account of the attempt by phase, scripts/warp_prof.py)
    const int tid = threadIdx.x;
#endif
    SbLane S;
    S.load(lane);
    const SbCosmo& CC = *(BATCH ? reinterpret_cast<const SbCosmo*>(sm + SB_SM_COSMO) : &A.c0);

    int item = 0, item_end = 0;
    if (A.items) { item = A.ibeg[blockIdx.x * SB_WARPS_PER_CTA + (threadIdx.x >> 5)]; item_end = A.ibeg[blockIdx.x * SB_WARPS_PER_CTA + (threadIdx.x >> 5) + 1]; }
#ifdef SB_TIME_PROF
    const unsigned long long tp0 = sb_now(); unsigned long long tpw = 0, tpi = tp0; int tpm = -1;
#define SB_TP_ITEM if (!GROUP && lane == 0 && tpm >= 0 && tpm < 8192) { const unsigned long long n_ = sb_now(); atomicAdd(&sb_tp_mode[tpm], n_ - tpi); tpi = n_; }
#else
#define SB_TP_ITEM
#endif
    while (true) {
        SB_TP_ITEM
        // work: either the atomic queue over whole modes, or this warp's static item list.  An item is a mode with an attempt
        // quota (> 0: park the mode after that many attempts and publish a continuation record) or a continuation (cont = 1:
        // wait for the record and integrate to the end).  Parking happens right after an accepted step, where the only live
        // state is (u, t, dt, controller, save index, table interval, counters): the resumed mode repeats exactly the
        // arithmetic an uninterrupted one would do, so results do not depend on the schedule.
        int mode, quota = 0, resume = 0;
        if (A.items) {
            if (item >= item_end) break;
            mode = A.items[3 * item]; quota = A.items[3 * item + 1]; resume = A.items[3 * item + 2];
            item++;
#ifdef SB_TIME_PROF
            tpm = mode; tpi = sb_now();
#endif
        } else if (GROUP) {
            __syncthreads(); // every lane is done with the previous mode (and with the exchange area)
            if (threadIdx.x == 0) reinterpret_cast<int*>(xch)[0] = atomicAdd(A.queue, 1);
            __syncthreads();
            const int qi = reinterpret_cast<const int*>(xch)[0];
            if (qi >= A.nk / A.G) break;
            mode = (A.order ? A.order[qi] : qi) * A.G + warp; // modes are laid out [group][lane]
        } else {
            int qi = 0;
            if (lane == 0) qi = atomicAdd(A.queue, 1);
            qi = __shfl_sync(SB_FULL, qi, 0);
            if (qi >= A.nk) break;
            mode = A.order ? A.order[qi] : qi;
#ifdef SB_TIME_PROF
            tpm = mode; tpi = sb_now();
#endif
        }
        if (BATCH) {
            __syncwarp();
            const double* src = reinterpret_cast<const double*>(A.cosmos + A.cosmo_of[mode]);
            if (lane < SB_COSMO_DOUBLES) sm[SB_SM_COSMO + lane] = src[lane];
            __syncwarp();
        }
        const double k = A.ks[mode];
        double t = A.tini[mode];
        const double tend = CC.tend;
        // GROUP: every decision that steers the step sequence is taken on COMMON data -- the smallest end time of the lanes and the primal
        // lane's save times (each lane interpolates at its own save times, which differ by O(δ), and, after the last common step, closes
        // with one private step to its own end time: its own "today")
        const SbCosmo* const C0 = GROUP ? A.cosmos + A.cosmo_of[mode - warp] : nullptr;
        const double tend0 = GROUP ? A.tendc : tend;
        const double* const saveat0 = GROUP ? C0->saveat : CC.saveat;
        bool closing = false; // GROUP: the private closing step of this lane
        long long naccept = 0, nreject = 0, nf = 0, nsolve = 0;
        int rc = SB_RC_SUCCESS, isave = 0;
        double* usave = A.usave ? A.usave + (size_t)mode * A.nsave * SB_N : nullptr;
        double* Sout = SRC ? A.S + (size_t)mode * A.nS * A.nsave : nullptr;
        int wstart = 0; // first valid slot of the current window of source values (see sb_source_flush)

        if (!(k > 0) || !isfinite(k)) { // reference: solve fails for k = 0 / NaN (test "Success checking", runtests.jl:358-361)
            for (int i = lane; i < SB_N; i += SB_WARP) A.uend[(size_t)mode * SB_N + i] = NAN;
            if (usave) for (int i = lane; i < A.nsave * SB_N; i += SB_WARP) usave[i] = NAN;
            if (SRC) for (int i = lane; i < A.nsave * A.nS; i += SB_WARP) Sout[i] = NAN;
            if (lane == 0) { A.retcode[mode] = SB_RC_UNSTABLE; for (int j = 0; j < 4; j++) A.stats[4 * mode + j] = 0; }
            continue;
        }
        if (lane < 7) kp[lane] = pow(k, (double)(lane - 3));
        const double lq0 = sb_ctl_lq0();
        double lqold = lq0; // log2 of the controller's qold
        double dt = 0;
        int jt = 0, it0 = 0;
        bool parked = false;
        if (resume) {
            int f = 0;
            if (lane == 0) { // bounded wait (5 s): a schedule whose lists are not all resident must fail, not hang
                unsigned long long t0, t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                while ((f = *(volatile int*)(A.flags + mode)) == 0) {
                    __nanosleep(200);
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 5000000000ull) { f = 3; break; }
                }
#ifdef SB_TIME_PROF
                tpw += sb_now() - t0;
#endif
            }
            f = __shfl_sync(SB_FULL, f, 0);
            __threadfence();
            if (f == 2) continue; // the mode finished within its first piece
            if (f == 3) {
                for (int i = lane; i < SB_N; i += SB_WARP) A.uend[(size_t)mode * SB_N + i] = NAN;
                if (lane == 0) { A.retcode[mode] = SB_RC_SCHEDULE; for (int j = 0; j < 4; j++) A.stats[4 * mode + j] = 0; }
                continue;
            }
            const double* c = A.cont + (size_t)mode * SB_CONT;
            for (int i = lane; i < SB_N; i += SB_WARP) u[i] = __ldcg(c + i);
            t = __ldcg(c + SB_N); dt = __ldcg(c + SB_N + 1); lqold = __ldcg(c + SB_N + 2);
            isave = (int)__ldcg(c + SB_N + 4); jt = (int)__ldcg(c + SB_N + 5); it0 = (int)__ldcg(c + SB_N + 6);
            wstart = isave & (SB_SWIN - 1); // the parking warp has stored the slots before isave of this window
            naccept = (long long)__ldcg(c + SB_N + 7); nreject = (long long)__ldcg(c + SB_N + 8); nf = (long long)__ldcg(c + SB_N + 9); nsolve = (long long)__ldcg(c + SB_N + 10);
            __syncwarp();
        } else {
        if (lane == 0) sb_initial_state(CC.spl, CC.P, t, k, U); // natural order
        __syncwarp();
        for (int i = lane; i < SB_N; i += SB_WARP) u[i] = U[sb_nat[i]]; // -> the integrator's path-contiguous order
        __syncwarp();
        while (isave < A.nsave && saveat0[isave] <= t) { // save points at (or before) the start
            const bool at = saveat0[isave] == t;
            if (usave) for (int i = lane; i < SB_N; i += SB_WARP) usave[(size_t)isave * SB_N + sb_nat[i]] = at ? u[i] : NAN;
            if (SRC) {
                double* so = sbuf + (isave & (SB_SWIN - 1));
                if (at) {
                    sb_source_point(S, CC.srcbg + (size_t)isave * SB_SRCBG_STRIDE, kp, k, t, CC.taurec, A.scale_k, A.nS, u, up, bs, bs + SB_NB, so, SB_SWIN, lane);
                } else if (lane < A.nS) so[lane * SB_SWIN] = NAN;
                sb_source_flush(sbuf, Sout, A.nS, A.nsave, isave, wstart, false, lane);
            }
            isave++;
        }
        }
        if (tend0 > t) {
            if (!resume) {
            jt = sb_interval(CC.tb, t); // knot interval of the current time
            jt = sb_basis_at(S, CC.tb, t, jt, kp, bs, bdv, lane);
            __syncwarp();
            sb_eval_f<false>(S, bs, u, f0, lane); nf++;
            sb_eval_dT(S, bs, bdv, u, dT, lane);
            { // automatic initial step (Hairer), order 5
                double d0 = 0, d1 = 0;
                for (int i = lane; i < SB_N; i += SB_WARP) { double sk = abstol + fabs(u[i]) * reltol; const double a0 = u[i] / sk, a1 = f0[i] / sk; d0 = fma(a0, a0, d0); d1 = fma(a1, a1, d1); } // (explicit fma chains: the split kernel forms the same sums)
                d0 = sqrt(warp_sum(d0) / SB_N); d1 = sqrt(warp_sum(d1) / SB_N);
                double dtmax = tend - t;
                double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                dt0 = fmin(dt0, dtmax);
                for (int i = lane; i < SB_N; i += SB_WARP) U[i] = u[i] + dt0 * f0[i];
                sb_basis_at(S, CC.tb, t + dt0, jt, kp, bs + SB_NB, nullptr, lane);
                __syncwarp();
                sb_eval_f<false>(S, bs + SB_NB, U, K, lane); nf++;
                double d2 = 0;
                for (int i = lane; i < SB_N; i += SB_WARP) { double sk = abstol + fabs(u[i]) * reltol; double r = (K[i] - f0[i]) / sk; d2 = fma(r, r, d2); }
                d2 = sqrt(warp_sum(d2) / SB_N) / dt0;
                double dm = fmax(d1, d2);
                double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2 + log10(dm)) / 5.0);
                dt = fmin(fmin(100 * dt0, dt1), dtmax);
                __syncwarp();
            }
            if (GROUP) { // one initial step for all lanes: the primal's
                if (threadIdx.x == 0) xch[1] = dt;
                __syncthreads();
                dt = xch[1];
            }
            }
            int jend = jt;       // interval of t + dt after the step (becomes jt on accept)
            bool moved = resume; // t advanced since f0, dT and the slot-0 basis were evaluated
            bool have0 = !resume; // SB_DEFER: slot 0 of the basis (value and derivative at t) is in shared memory (carried over from the last stage time of the accepted step)
            for (int it = it0;; it++) {
                if (it >= A.maxiters) { rc = SB_RC_MAXITERS; break; }
                if (quota > 0 && moved && it - it0 >= quota) { // park: publish the continuation record
                    double* c = A.cont + (size_t)mode * SB_CONT;
                    for (int i = lane; i < SB_N; i += SB_WARP) c[i] = u[i];
                    if (lane == 0) {
                        c[SB_N] = t; c[SB_N + 1] = dt; c[SB_N + 2] = lqold; c[SB_N + 3] = 0; c[SB_N + 4] = isave; c[SB_N + 5] = jt; c[SB_N + 6] = it;
                        c[SB_N + 7] = (double)naccept; c[SB_N + 8] = (double)nreject; c[SB_N + 9] = (double)nf; c[SB_N + 10] = (double)nsolve;
                    }
                    if (SRC && (isave & (SB_SWIN - 1)) != 0) sb_source_flush(sbuf, Sout, A.nS, A.nsave, isave - 1, wstart, true, lane); // partial window of source values
                    parked = true;
                    break;
                }
                SB_WPROF(0)
                bool last = false;
                double dtc = dt; // the step as the (shared) controller sees it; dt is what this lane integrates over (differs only on the last step of a GROUP lane)
                if (GROUP && closing) last = true; // dt = tend − t was set when the lockstep phase ended
                else if (t + dt >= tend0 - 100 * 2.2e-16 * fabs(tend0)) { dtc = tend0 - t; dt = GROUP ? dtc : tend - t; last = true; }
                // basis at the stage times of this attempt (one batched table look-up)
#if SB_DEFER
                if (moved && !have0) { jt = sb_basis_slot0(S, CC.tb, t, jt, kp, bs, bdv, kp + 8, lane); have0 = true; } // (a resumed mode: slot 0 is not in shared memory)
                sb_split_basis_issue(CC.tb, t, dt, jt, kp + 8, tabs, reinterpret_cast<int*>(kp + 8), lane SB_WPROF_PASS); // rows of the stage times 2..6: in flight during the factorisation
                jend = reinterpret_cast<const int*>(kp + 8)[0];
#elif SB_TMA
                jend = sb_basis_batch(S, CC.tb, t, dt, jt, moved, kp, bs, bdv, kp + 8, lane, tabs, &tpar);
#else
                jend = sb_basis_batch(S, CC.tb, t, dt, jt, moved, kp, bs, bdv, kp + 8, lane);
#endif
                __syncwarp();
                SB_WPROF(1)
                if (moved) { sb_eval_f<false>(S, bs, u, f0, lane); nf++; sb_eval_dT(S, bs, bdv, u, dT, lane); moved = false; }
                SB_WPROF(2)
                // factor W = I/(γ dt) − J(t) = B − p φᵀ − q ψᵀ
                sb_factor(S, sb_wdiag(dt), bs, di, up, mm, blk, lane);
                SB_WPROF(3)
#pragma unroll
                for (int r = 0; r < SB_R; r++) {
                    const int i = r * 32 + lane;
                    if (i < SB_N) { Zp[i] = S.pqc[2 * r] * bs[SB_LO16(S.pqi[r])]; Zq[i] = S.pqc[2 * r + 1] * bs[SB_HI16(S.pqi[r])]; }
                }
#if SB_Z3
                { // the stage-1 right-hand side f0 + dt·d1·dT does not depend on Z: its solve rides along with the two Woodbury columns
                    const double hd0 = dt * cd[0];
#pragma unroll
                    for (int r = 0; r < SB_R; r++) { const int i = r * 32 + lane; if (i < SB_N) K[i] = f0[i] + hd0 * dT[i]; }
                }
                __syncwarp();
                { double* const zz[3] = {Zp, Zq, K}; sb_bsolve<3>(S, zz, di, up, mm, blk, lane); }
                nsolve += 3;
#else
                __syncwarp();
                { double* const zz[2] = {Zp, Zq}; sb_bsolve<2>(S, zz, di, up, mm, blk, lane); }
                nsolve += 2;
#endif
                double m11, m12, m21, m22;
                sb_hub_dots(S, bs, Zp, m11, m21, lane);
                sb_hub_dots(S, bs, Zq, m12, m22, lane);
                m11 = 1 - m11; m12 = -m12; m21 = -m21; m22 = 1 - m22;
                const double idet = sb_rcp(m11 * m22 - m12 * m21);
                const double idt = sb_rcp(dt);
                SB_WPROF(4)
#if SB_DEFER
                sb_split_basis_sweep(S, kp, bs, sm + SB_SM_BD5, kp + 8, tabs, tpar, lane, 0, 1); // the rows have arrived long ago: wait, Hermite sweep of slots 1..5
                tpar ^= 1u;
                __syncwarp();
#endif
                SB_WPROF(5)
                // 8 stages.  Per stage: (A) lane-local: finish k_{s-1} with its pending Woodbury correction and accumulate U_s = u + Σ a_sj k_j,
                // R_s = Σ (C_sj/dt) k_j; (B) f(U_s) with the right-hand side assembled in the same pass; (C) B-solve.  The hub dot products of the
                // rank-2 correction k_s += Z c and the correction itself are applied lazily in (A) of the next stage: no extra pass, no extra barrier.
                double c1p = 0, c2p = 0;
#if SB_STAGE_UNROLL
#pragma unroll
#endif
                for (int s = 0; s < 8; s++) {
                    double* ks = K + s * SB_N;
                    const double hd_ = dt * cd[s];
                    if (s > 0) {
                        // U_s = u + Σ_{j<s} a_sj k_j (stages 7, 8: U_s = U_{s-1} + k_{s-1}, the tableau rows repeat) and R_s = Σ_{j<s} (C_sj/dt) k_j;
                        // absent terms are skipped by warp-uniform branches.
                        // The Woodbury coefficients of k_{s-1} need two warp reductions (hub dot products of the solve result); their
                        // butterfly levels are interleaved with the accumulation of the older stages, which does not depend on them.
                        double Racc[SB_R], ua[SB_R];
                        int ii[SB_R];
                        const bool inc = s >= 6;
                        const int nold = s - 1;
                        double* kprev = ks - SB_N;
                        double h1 = 0, h2 = 0, t1, t2;
#pragma unroll
                        for (int t = 0; t < SB_TPH; t++) h1 += S.phc[t] * bs[SB_HI16(S.phi[t])] * kprev[SB_LO16(S.phi[t])];
#pragma unroll
                        for (int t = 0; t < SB_TPS; t++) h2 += S.psc[t] * bs[SB_HI16(S.psi[t])] * kprev[SB_LO16(S.psi[t])];
#if SB_STAGE_BF
#define SB_STAGE_TERM(j, WITHU)                                                                                   \
    {                                                                                                             \
        const bool on_ = nold > (j);                                                                              \
        const double a_ = inc ? 0.0 : cA[s][j], c_ = cC[s][j] * idt;                                              \
        _Pragma("unroll") for (int r = 0; r < SB_R; r++) {                                                        \
            const double kl = K[(j) * SB_N + ii[r]];                                                              \
            const double kj = on_ ? kl : 0.0;                                                                     \
            if (WITHU) ua[r] = fma(a_, kj, ua[r]);                                                                \
            Racc[r] = fma(c_, kj, Racc[r]);                                                                       \
        }                                                                                                         \
    }
#else
#define SB_STAGE_TERM(j, WITHU)                                                                                   \
    if (nold > (j)) {                                                                                             \
        const double a_ = cA[s][j], c_ = cC[s][j] * idt;                                                          \
        _Pragma("unroll") for (int r = 0; r < SB_R; r++) {                                                        \
            const double kj = K[(j) * SB_N + ii[r]];                                                              \
            if (WITHU && !inc) ua[r] = fma(a_, kj, ua[r]);                                                        \
            Racc[r] = fma(c_, kj, Racc[r]);                                                                       \
        }                                                                                                         \
    }
#endif
#define SB_LEVEL_ISSUE(o) t1 = __shfl_xor_sync(SB_FULL, h1, o); t2 = __shfl_xor_sync(SB_FULL, h2, o);
#define SB_LEVEL_ADD h1 += t1; h2 += t2;
                        SB_LEVEL_ISSUE(16)
#pragma unroll
                        for (int r = 0; r < SB_R; r++) { const int i = min(r * 32 + lane, SB_N - 1); ii[r] = i; ua[r] = inc ? U[i] : u[i]; Racc[r] = 0; }
                        SB_LEVEL_ADD SB_LEVEL_ISSUE(8)
                        SB_STAGE_TERM(0, true)
                        SB_LEVEL_ADD SB_LEVEL_ISSUE(4)
                        SB_STAGE_TERM(1, true)
                        SB_LEVEL_ADD SB_LEVEL_ISSUE(2)
                        SB_STAGE_TERM(2, true)
                        SB_LEVEL_ADD SB_LEVEL_ISSUE(1)
                        SB_STAGE_TERM(3, true)
                        SB_STAGE_TERM(4, false)
                        SB_STAGE_TERM(5, false)
                        SB_LEVEL_ADD
                        c1p = (m22 * h1 - m12 * h2) * idet; c2p = (-m21 * h1 + m11 * h2) * idet;
                        {
                            const double al = cA[s][s - 1], cl = cC[s][s - 1] * idt;
#pragma unroll
                            for (int r = 0; r < SB_R; r++) {
                                const int i = ii[r];
                                const double kp_ = kprev[i] + Zp[i] * c1p + Zq[i] * c2p;
                                ua[r] = fma(al, kp_, ua[r]);
                                Racc[r] = fma(cl, kp_, Racc[r]);
                                if (r * 32 + lane < SB_N) kprev[i] = kp_;
                            }
                        }
#undef SB_LEVEL_ISSUE
#undef SB_LEVEL_ADD
#undef SB_STAGE_TERM
#pragma unroll
                        for (int r = 0; r < SB_R; r++) if (r * 32 + lane < SB_N) U[ii[r]] = ua[r];
                        __syncwarp();
                        SB_WPROF(6)
                        sb_eval_f<true>(S, bs + cslot[s] * SB_NB, U, ks, lane, Racc, hd_, dT); nf++;
                        SB_WPROF(7)
                    } else {
#if SB_Z3
                        continue; // solved together with Z
#else
#pragma unroll
                        for (int r = 0; r < SB_R; r++) { const int i = r * 32 + lane; if (i < SB_N) ks[i] = f0[i] + hd_ * dT[i]; }
                        __syncwarp();
#endif
                    }
                    { double* const k1[1] = {ks}; sb_bsolve<1>(S, k1, di, up, mm, blk, lane); } nsolve++;
                    SB_WPROF(8)
                }
                {
                    double s1, s2;
                    sb_hub_dots(S, bs, K + 7 * SB_N, s1, s2, lane);
                    c1p = (m22 * s1 - m12 * s2) * idet; c2p = (-m21 * s1 + m11 * s2) * idet;
                }
#pragma unroll
                for (int r = 0; r < SB_R; r++) { const int i = r * 32 + lane; if (i < SB_N) K[7 * SB_N + i] = sb_wcorr(K[7 * SB_N + i], Zp[i], c1p, Zq[i], c2p); }
                __syncwarp();
                // error estimate: k8 (Rodas5P), RMS norm scaled by abstol + reltol·max(|u|,|unew|)
                double es = 0; bool bad = false;
                if (GROUP && !closing) {
                    // publish (u_n, u_{n+1}, k_8) of this lane, then every warp forms the SAME norm over primal and partials:
                    // r_i = Dual(k8_i) / (abstol + reltol·max(‖u_n,i‖, ‖u_{n+1},i‖)), ‖x‖² = x² + Σ_j (∂_j x)², EEst² = Σ_i ‖r_i‖² / N
                    const int G = A.G;
                    double* X = xch + SB_XCH_HDR + (size_t)(((it & 1) * G + warp) * 3) * SB_N;
                    for (int i = lane; i < SB_N; i += SB_WARP) { const double k8 = K[7 * SB_N + i]; X[i] = u[i]; X[SB_N + i] = U[i] + k8; X[2 * SB_N + i] = k8; }
                    __syncthreads();
                    const double* X0 = xch + SB_XCH_HDR + (size_t)((it & 1) * G * 3) * SB_N;
                    for (int i = lane; i < SB_N; i += SB_WARP) {
                        const double u0 = X0[i], u1 = X0[SB_N + i], k8 = X0[2 * SB_N + i];
                        double s0 = u0 * u0, s1 = u1 * u1, sk = k8 * k8;
                        for (int j = 1; j < G; j++) {
                            const double* Xj = X0 + (size_t)j * 3 * SB_N;
                            const double d = A.invdelta[j];
                            const double a0 = (Xj[i] - u0) * d, a1 = (Xj[SB_N + i] - u1) * d, ak = (Xj[2 * SB_N + i] - k8) * d;
                            s0 += a0 * a0; s1 += a1 * a1; sk += ak * ak;
                        }
                        const double sc = abstol + reltol * fmax(sqrt(s0), sqrt(s1));
#if SB_NORMRCP
                        const double r = sqrt(sk) * sb_rcp(sc); // with vanishing partials this is |k8| · rcp(abstol + reltol·max(|u_n|, |u_{n+1}|)): the plain norm, bit for bit
#else
                        const double r = sqrt(sk) / sc;
#endif
                        es = fma(r, r, es);
                    }
                } else
#if SB_NORMRCP
                for (int i = lane; i < SB_N; i += SB_WARP) { double k8 = K[7 * SB_N + i], un = U[i] + k8; double r = k8 * sb_rcp(abstol + reltol * fmax(fabs(u[i]), fabs(un))); es = fma(r, r, es); }
#else
                for (int i = lane; i < SB_N; i += SB_WARP) { double k8 = K[7 * SB_N + i], un = U[i] + k8; double r = k8 / (abstol + reltol * fmax(fabs(u[i]), fabs(un))); es = fma(r, r, es); }
#endif
                double EEst = sb_rms(warp_sum(es));
                SB_WPROF(9)
                if (A.trace && mode == 0 && lane == 0 && it < A.ntrace) { A.trace[3 * it] = t; A.trace[3 * it + 1] = dt; A.trace[3 * it + 2] = EEst; }
                if (GROUP && closing) { if (!isfinite(EEst)) { rc = SB_RC_UNSTABLE; break; } EEst = fmin(EEst, 1.0); } // the closing step (O(δ) long) is taken as it is
                if (!isfinite(EEst)) { nreject++; dt = dtc / 5; if (dt < 1e-15 * fabs(t)) { rc = SB_RC_UNSTABLE; break; } continue; }
                const double lE = log2(EEst); // (−inf for EEst = 0: q = 1/qmax then, and qold falls back to its floor)
                const double q = (EEst == 0.0) ? 0.1 : sb_ctl_qacc(sb_ctl_q(lE, lqold));
                if (EEst > 1) { nreject++; dt = dtc * sb_rcp(sb_ctl_qrej(sb_ctl_q11(lE))); if (dt < 1e-15 * fabs(t)) { rc = SB_RC_DTMIN; break; } continue; }
                naccept++;
                const double dtnew = dtc * sb_rcp(q);
                lqold = fmax(lE, lq0); // qold = max(EEst, 1e-4)
                SB_WPROF(10)
                const bool common = GROUP && !closing;                                   // lockstep phase: decisions on common data
                const double tn = last ? (common ? tend0 : tend) : t + dt;
                const double tn0 = common ? (last ? tend0 : t + dtc) : tn;
                const double* const svd = common ? saveat0 : CC.saveat;
                if (isave < A.nsave && svd[isave] <= tn0) { // dense output (4th order), vectors stored over dT, f0, Zp
                    for (int i = lane; i < SB_N; i += SB_WARP) {
                        double a1 = 0, a2 = 0, a3 = 0;
                        for (int j = 0; j < 8; j++) { double kj = K[j * SB_N + i]; a1 = fma(cH[0][j], kj, a1); a2 = fma(cH[1][j], kj, a2); a3 = fma(cH[2][j], kj, a3); }
                        dT[i] = a1; f0[i] = a2; Zp[i] = a3;
                    }
                    while (isave < A.nsave && svd[isave] <= tn0) {
                        const double ts = CC.saveat[isave];
                        const double th = (ts - t) / dt, t1 = 1 - th;
                        const bool atend = svd[isave] == tn0; // a save time that IS the step end takes the step result
                        for (int i = lane; i < SB_N; i += SB_WARP) {
                            const double un = U[i] + K[7 * SB_N + i];
                            const double v = atend ? un : sb_dense(th, t1, u[i], un, dT[i], f0[i], Zp[i]);
                            if (usave) usave[(size_t)isave * SB_N + sb_nat[i]] = v;
                            if (SRC) di[i] = v;
                        }
                        if (SRC) { // the sources at this save time, from the state just formed (reference: output_func of solvept, fourier.jl:272-278)
                            __syncwarp();
                            sb_source_point(S, CC.srcbg + (size_t)isave * SB_SRCBG_STRIDE, kp, k, ts, CC.taurec, A.scale_k, A.nS, di, up, bs, bs + SB_NB, sbuf + (isave & (SB_SWIN - 1)), SB_SWIN, lane);
                            sb_source_flush(sbuf, Sout, A.nS, A.nsave, isave, wstart, false, lane);
                        }
                        isave++;
                    }
                }
#if SB_DEFER
                for (int m = lane; m < SB_NB; m += SB_WARP) { bs[m] = bs[5 * SB_NB + m]; bdv[m] = sm[SB_SM_BD5 + m]; } // the basis at the new time is the one of the last stage time (c = 1)
#endif
                for (int i = lane; i < SB_N; i += SB_WARP) { double un = U[i] + K[7 * SB_N + i]; if (isnan(un)) bad = true; u[i] = un; }
                t = tn;
                bad = __any_sync(SB_FULL, bad);
                __syncwarp();
                if (bad) { rc = SB_RC_UNSTABLE; break; }
                if (last) {
                    if (GROUP && !closing && tend > t) { closing = true; dt = tend - t; jt = jend; moved = true; continue; } // private closing step to this lane's own end time
                    break;
                }
                dt = dtnew;
                jt = jend;
                moved = true;
                SB_WPROF(11)
            }
        }
#ifdef SB_WARP_PROF
        if (!BATCH && !GROUP) { SB_PROF_OUT }
#endif
        if (parked) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(A.flags + mode, 1);
            continue;
        }
        for (int i = lane; i < SB_N; i += SB_WARP) A.uend[(size_t)mode * SB_N + sb_nat[i]] = u[i];
        if (GROUP && rc == SB_RC_SUCCESS) // the lane whose end time IS the common one has no closing step: its save times at the end are due now
            while (isave < A.nsave && CC.saveat[isave] <= t) {
                if (usave) for (int i = lane; i < SB_N; i += SB_WARP) usave[(size_t)isave * SB_N + sb_nat[i]] = u[i];
                if (SRC) {
                    sb_source_point(S, CC.srcbg + (size_t)isave * SB_SRCBG_STRIDE, kp, k, t, CC.taurec, A.scale_k, A.nS, u, up, bs, bs + SB_NB, sbuf + (isave & (SB_SWIN - 1)), SB_SWIN, lane);
                    sb_source_flush(sbuf, Sout, A.nS, A.nsave, isave, wstart, false, lane);
                }
                isave++;
            }
        for (; isave < A.nsave; isave++) { // save times the mode never reached (failed solve)
            if (usave) for (int i = lane; i < SB_N; i += SB_WARP) usave[(size_t)isave * SB_N + i] = NAN;
            if (SRC) { if (lane < A.nS) sbuf[lane * SB_SWIN + (isave & (SB_SWIN - 1))] = NAN; sb_source_flush(sbuf, Sout, A.nS, A.nsave, isave, wstart, false, lane); }
        }
        if (lane == 0) { A.retcode[mode] = rc; A.stats[4 * mode] = naccept; A.stats[4 * mode + 1] = nreject; A.stats[4 * mode + 2] = nf; A.stats[4 * mode + 3] = nsolve; }
        if (quota > 0 && lane == 0) atomicExch(A.flags + mode, 2);
        __syncwarp();
    }
#ifdef SB_TIME_PROF
    if (!GROUP && lane == 0) { const int w_ = blockIdx.x * SB_WARPS_PER_CTA + (threadIdx.x >> 5); if (w_ < 4096) { sb_tp_list[3 * w_] = tp0; sb_tp_list[3 * w_ + 1] = sb_now(); sb_tp_list[3 * w_ + 2] = tpw; } }
#endif
}

// ================================================================================================ split kernel
// One CTA of SB_R warps per mode ("one CTA per k-mode"): for launches with fewer modes than SMs' worth of warps (BASELINE config 1: 100 modes;
// the reference's default C_l path: 61) the warp-per-mode kernel leaves three of four sub-partitions idle and its attempt is a single
// warp's dependent instruction stream.  Here warp w owns round rb of the rows (32 rows each): the row-parallel phases -- basis sweep,
// Jacobian scatter, f-evaluations, stage combinations, error norm, dense output -- shrink to one round per warp, the three columns of the
// first solve go to three warps, and the phases that live on lanes rather than rows (path recurrences, 5×5 block, single-column solves) stay
// on warp 0, with CTA barriers at the phase boundaries.  Every warp carries the scalar control state and takes identical decisions from
// shared data; sums over rows are formed per lane over the rounds in the single-warp order before the butterfly, so the results are
// bit-identical to sb_integrate_kernel<false> (tests/test_gpu_parity.py::test_split_kernel_is_bit_identical).
#if SB_TMA && SB_BSLOT && SB_R >= 2 && SB_R <= 4 && SB_TR == 0 && SB_NTOP == 1 && SB_UNIQUE_TARGETS
#define SB_HAVE_SPLIT 1
static_assert(SB_J_F2 < 32 && SB_J_G0 < 32 && SB_J_G2 < 32, "the split kernel closes the source evaluation on the warp of round 0");
#define SB_XS_SUM 8 // exchange area: ints [0] work index, [1] knot interval of t + dt; doubles [2..5] the controller warp's mailbox; [8 + r·32 + lane] per-lane terms of round r
#define SB_XS_P1 2  // answer: q11 / qold^β2 (accepted) or q11 (rejected) of the current command
#define SB_XS_CMD 3 // command to the controller warp: the error estimate (NaN: exit)
#define SB_XS_P2 4  // [4] log2 qold posted with the command, [5] log2 EEst returned with the answer
#define SB_XS_KC (SB_XS_SUM + 32 * SB_R)   // corrected stage vectors k_1..k_7, [7][SB_N]
#define SB_XS_BD5 (SB_XS_KC + 7 * SB_N)    // derivative basis at the last stage time, [SB_NB]
#define SB_XS_DOUBLES (SB_XS_BD5 + SB_NB)
// Σ_rows v_row² where v is this thread's row value (0 for an absent row): per lane the fma chain over the rounds in the single-warp order, then the butterfly
__device__ __forceinline__ double sb_split_sumsq(double v, double* xs, int rb, int lane) {
    xs[SB_XS_SUM + rb * 32 + lane] = v;
    sb_rows_sync();
    double t = 0;
#pragma unroll
    for (int r = 0; r < SB_R; r++) { const double x = xs[SB_XS_SUM + r * 32 + lane]; t = fma(x, x, t); }
    sb_rows_sync();
    return warp_sum(t);
}
__device__ __forceinline__ void sb_split_bar_others() { asm volatile("bar.sync 1, %0;" ::"n"(SB_WARP * (SB_R - 1)) : "memory"); }
__global__ void __launch_bounds__(SB_SPLIT_THREADS, 1) sb_integrate_split_kernel(const __grid_constant__ SbSolveArgs A) {
    constexpr bool SPL = true;
    extern __shared__ __align__(16) double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    constexpr int NT = SB_SPLIT_NT; // row threads
    const int rb = SB_R - 1 - warp; // warp 0 also runs the serial phases: it takes the last (shortest, narrowest) round
    const bool w0 = warp == 0;
    double* const sm = sm_all;
    double *u = sm + SB_SM_U, *U = sm + SB_SM_UU, *K = sm + SB_SM_K, *dT = sm + SB_SM_DT, *f0 = sm + SB_SM_F0, *di = sm + SB_SM_DI, *up = sm + SB_SM_UP,
           *mm = sm + SB_SM_MM, *Zp = sm + SB_SM_ZP, *Zq = sm + SB_SM_ZQ, *bs = sm + SB_SM_BS, *bdv = sm + SB_SM_BD, *blk = sm + SB_SM_BLK, *kp = sm + SB_SM_KP;
    double* const sbuf = sm + SB_SM_SBUF;
    double* const tabs = sm + SB_SM_TAB;
    double* const xs = sm + SB_SM_DOUBLES;
    unsigned tpar = 0;
    SB_PROF_DECL
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb_smem_u32(sm + SB_SM_MBAR)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (warp == SB_R) {
        // Controller warp.  The step controller needs log2 EEst and an exp2 (sb_ctl_q) -- ≈0.6 k cycles of one warp's time -- before the next step size is
        // known.  The row warps post (EEst, log2 qold) as soon as the error norm is formed (barrier 3) and go on with everything of an accepted step that does
        // not depend on the new step size (dense output, state update, f0 and dT at the new time); they pick the answer up at barrier 4.  NaN is the exit command.
        while (true) {
            asm volatile("bar.sync 3, %0;" ::"n"(SB_SPLIT_THREADS) : "memory");
            const double E = xs[SB_XS_CMD], lq = xs[SB_XS_P2];
            if (E != E) break;
            const double l = log2(E);
            const double r = (E > 1) ? sb_ctl_q11(l) : sb_ctl_q(l, lq);
            if (lane == 0) { xs[SB_XS_P1] = r; xs[SB_XS_P2 + 1] = l; }
            __threadfence_block();
            asm volatile("bar.arrive 4, %0;" ::"n"(SB_SPLIT_THREADS) : "memory");
        }
        return;
    }
    const double reltol = A.reltol, abstol = A.abstol;
    const bool SRC = A.S != nullptr;
    SbLane S;
    S.load_split(lane, rb);
    auto ctl_send = [&](double E, double lq) {
        if (tid == 0) { xs[SB_XS_CMD] = E; xs[SB_XS_P2] = lq; }
        __threadfence_block();
        asm volatile("bar.arrive 3, %0;" ::"n"(SB_SPLIT_THREADS) : "memory");
    };
    const SbCosmo& CC = A.c0;
    const int i0 = rb * 32 + lane;         // this thread's row
    const bool has = i0 < SB_N;
    const int ic = has ? i0 : SB_N - 1;    // clamped: loads of absent rows read the last row (results discarded)

    while (true) {
        sb_rows_sync(); // every warp is done with the previous mode
        if (tid == 0) reinterpret_cast<int*>(xs)[0] = atomicAdd(A.queue, 1);
        sb_rows_sync();
        const int qi = reinterpret_cast<const int*>(xs)[0];
        if (qi >= A.nk) { ctl_send(NAN, 0.0); break; }
        const int mode = A.order ? A.order[qi] : qi;
        const double k = A.ks[mode];
        double t = A.tini[mode];
        const double tend = CC.tend;
        long long naccept = 0, nreject = 0, nf = 0, nsolve = 0;
        int rc = SB_RC_SUCCESS, isave = 0, wstart = 0;
        double* usave = A.usave ? A.usave + (size_t)mode * A.nsave * SB_N : nullptr;
        double* Sout = SRC ? A.S + (size_t)mode * A.nS * A.nsave : nullptr;
        if (!(k > 0) || !isfinite(k)) {
            for (int i = tid; i < SB_N; i += NT) A.uend[(size_t)mode * SB_N + i] = NAN;
            if (usave) for (int i = tid; i < A.nsave * SB_N; i += NT) usave[i] = NAN;
            if (SRC) for (int i = tid; i < A.nsave * A.nS; i += NT) Sout[i] = NAN;
            if (tid == 0) { A.retcode[mode] = SB_RC_UNSTABLE; for (int j = 0; j < 4; j++) A.stats[4 * mode + j] = 0; }
            continue;
        }
        if (tid < 7) kp[tid] = pow(k, (double)(tid - 3));
        const double lq0 = sb_ctl_lq0();
        double lqold = lq0; // log2 of the controller's qold
        double dt = 0;
        int jt = 0;
        if (tid == 0) sb_initial_state(CC.spl, CC.P, t, k, U); // natural order
        if (has) mm[i0] = 0; // multiplier slots: cleared once per mode (and after a non-finite attempt), see sb_factor
        for (int i = tid; i < SB_TOPSTORE; i += NT) blk[i] = 0;
        sb_rows_sync();
        if (has) u[i0] = U[sb_nat[i0]];
        sb_rows_sync();
        // one save point from the state in `st` (all warps; st complete and visible on entry)
        auto emit = [&](const double* st, double ts, bool valid) {
            if (SRC) {
                double* so = sbuf + (isave & (SB_SWIN - 1));
                if (valid) {
                    const double* sb = CC.srcbg + (size_t)isave * SB_SRCBG_STRIDE;
                    for (int m = tid; m < SB_NB; m += NT) {
                        const unsigned pk = sb_basis_pack[m];
                        const double kk = kp[SB_HI16(pk)];
                        bs[m] = kk * __ldg(sb + 8 + SB_LO16(pk));
                        bs[SB_NB + m] = kk * __ldg(sb + 8 + SB_NBETA + SB_LO16(pk));
                    }
                    sb_rows_sync();
                    double hub[2];
                    sb_eval_f<false, SPL>(S, bs, st, up, lane, nullptr, 0.0, nullptr, hub, rb); // u̇ into `up`
                    sb_rows_sync();
                    if (rb == 0) { // the rows of Π = F2 + G0 + G2 live in round 0; its warp closes the evaluation (same sums, same order as sb_source_point)
                        const double* b = bs; const double* bd = bs + SB_NB;
                        double a = 0, pdd = 0;
#pragma unroll
                        for (int tt = 0; tt < SB_TPS; tt++) a = sb_dd_term(S.psc[tt], bd[SB_HI16(S.psi[tt])], st[SB_LO16(S.psi[tt])], b[SB_HI16(S.psi[tt])], up[SB_LO16(S.psi[tt])], a);
                        if (i0 == SB_J_F2 || i0 == SB_J_G0 || i0 == SB_J_G2) {
#pragma unroll
                            for (int w = 0; w < SB_WMAX; w++) pdd = sb_dd_term(S.ec[w], bd[SB_E_B(S.ei[w])], st[SB_E_COL(S.ei[w])], b[SB_E_B(S.ei[w])], up[SB_E_COL(S.ei[w])], pdd);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(SB_FULL, a, o); pdd += __shfl_xor_sync(SB_FULL, pdd, o); }
                        if (lane == 0)
                            sb_source_close(sb, k, ts, CC.taurec, A.scale_k, A.nS, hub[0], hub[1], a, st[SB_J_F2] + st[SB_J_G0] + st[SB_J_G2], up[SB_J_F2] + up[SB_J_G0] + up[SB_J_G2], pdd, st[SB_J_TB],
                                            up[SB_J_TB], st[SB_J_F0], st[SB_J_PHI], so, SB_SWIN);
                    }
                } else if (tid < A.nS) so[tid * SB_SWIN] = NAN;
                sb_rows_sync();
                if (w0) sb_source_flush(sbuf, Sout, A.nS, A.nsave, isave, wstart, false, lane);
                sb_rows_sync();
            }
            isave++;
        };
        while (isave < A.nsave && CC.saveat[isave] <= t) { // save points at (or before) the start
            const bool at = CC.saveat[isave] == t;
            if (usave && has) usave[(size_t)isave * SB_N + sb_nat[i0]] = at ? u[i0] : NAN;
            emit(u, t, at);
        }
        if (tend > t) {
            jt = sb_interval(CC.tb, t);
            jt = sb_basis_at(S, CC.tb, t, jt, kp, bs, bdv, lane); // (every warp: identical values into the same words)
            sb_rows_sync();
            sb_eval_f<false, SPL>(S, bs, u, f0, lane, nullptr, 0.0, nullptr, nullptr, rb); nf++;
            sb_eval_dT<SPL>(S, bs, bdv, u, dT, lane, rb);
            sb_rows_sync();
            { // automatic initial step (Hairer), order 5 -- sums over rows in the single-warp order
                double d0 = 0, d1 = 0;
                if (has) { const double sk = abstol + fabs(u[i0]) * reltol; d0 = u[i0] / sk; d1 = f0[i0] / sk; }
                d0 = sqrt(sb_split_sumsq(d0, xs, rb, lane) / SB_N); d1 = sqrt(sb_split_sumsq(d1, xs, rb, lane) / SB_N);
                const double dtmax = tend - t;
                double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                dt0 = fmin(dt0, dtmax);
                if (has) U[i0] = u[i0] + dt0 * f0[i0];
                sb_basis_at(S, CC.tb, t + dt0, jt, kp, bs + SB_NB, nullptr, lane);
                sb_rows_sync();
                sb_eval_f<false, SPL>(S, bs + SB_NB, U, K, lane, nullptr, 0.0, nullptr, nullptr, rb); nf++;
                sb_rows_sync();
                double d2 = 0;
                if (has) { const double sk = abstol + fabs(u[i0]) * reltol; d2 = (K[i0] - f0[i0]) / sk; }
                d2 = sqrt(sb_split_sumsq(d2, xs, rb, lane) / SB_N) / dt0;
                const double dm = fmax(d1, d2);
                const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2 + log10(dm)) / 5.0);
                dt = fmin(fmin(100 * dt0, dt1), dtmax);
            }
            int jend = jt;
            bool moved = false;
            bool pend = false;  // the controller's answer to the last accepted step is still outstanding (dt holds that step's size meanwhile)
            double Epend = 0;   // its error estimate
            // the controller's answer: the new step size after an accepted (dt / q) or a rejected step, the same expressions as sb_integrate_kernel
            auto ctl_resolve = [&](double E, bool rejected, double dtstep) {
                asm volatile("bar.sync 4, %0;" ::"n"(SB_SPLIT_THREADS) : "memory");
                const double r = xs[SB_XS_P1], lE = xs[SB_XS_P2 + 1];
                if (rejected) return dtstep * sb_rcp(sb_ctl_qrej(r));
                const double q = (E == 0.0) ? 0.1 : sb_ctl_qacc(r);
                lqold = fmax(lE, lq0); // qold = max(EEst, 1e-4)
                return dtstep * sb_rcp(q);
            };
            for (int it = 0;; it++) {
                if (it >= A.maxiters) { rc = SB_RC_MAXITERS; break; }
                SB_PROF(0)
                sb_rows_sync(); // the previous attempt's readers of bs / the staging area are done; u and slot 0 of the basis (time t) are in place
                SB_PROF(1)
                if (moved) {
                    sb_eval_f<false, SPL>(S, bs, u, f0, lane, nullptr, 0.0, nullptr, nullptr, rb); nf++;
                    sb_eval_dT<SPL>(S, bs, bdv, u, dT, lane, rb);
                    moved = false; // (f0, dT are only read on the thread's own row)
                }
                SB_PROF(2)
                if (pend) { dt = ctl_resolve(Epend, false, dt); pend = false; }
                bool last = false;
                if (t + dt >= tend - 100 * 2.2e-16 * fabs(tend)) { dt = tend - t; last = true; }
                SB_PROF(10)
                // factor W = I/(γ dt) − J(t) = B − p φᵀ − q ψᵀ: rows by all warps, eliminations by warp 0.  Meanwhile the other warps set up Z and the
                // stage-1 right-hand side and bring in the basis at the stage times 2..6 (table rows by TMA, Hermite sweep): none of it is on warp 0's path.
                sb_factor<SPL>(S, sb_wdiag(dt), bs, di, up, mm, blk, lane, rb, w0 SB_PROF_PASS);
                if (has) {
                    Zp[i0] = S.pqc[0] * bs[SB_LO16(S.pqi[0])]; Zq[i0] = S.pqc[1] * bs[SB_HI16(S.pqi[0])];
                    K[i0] = f0[i0] + dt * cd[0] * dT[i0];
                }
                if (!w0) {
                    if (warp == 1) sb_split_basis_issue(CC.tb, t, dt, jt, kp + 8, tabs, reinterpret_cast<int*>(xs) + 1, lane);
                    sb_split_bar_others(); // slot descriptors visible to the sweeping warps
                    sb_split_basis_sweep(S, kp, bs, xs + SB_XS_BD5, kp + 8, tabs, tpar, lane, warp - 1, SB_R - 1);
                }
                tpar ^= 1u;
                sb_rows_sync();
                jend = reinterpret_cast<const int*>(xs)[1];
                SB_PROF(3)
                // the three columns of the first solve: one per warp (with fewer than three warps, warp 0 takes the rest)
                {
                    double* const cols[3] = {Zp, Zq, K};
#pragma unroll
                    for (int c = 0; c < 3; c++) if ((c < SB_R ? c : 0) == warp) { double* const z1[1] = {cols[c]}; sb_bsolve<1>(S, z1, di, up, mm, blk, lane); }
                }
                nsolve += 3;
                sb_rows_sync();
                SB_PROF(4)
                double m11, m12, m21, m22;
                sb_hub_dots(S, bs, Zp, m11, m21, lane);
                sb_hub_dots(S, bs, Zq, m12, m22, lane);
                m11 = 1 - m11; m12 = -m12; m21 = -m21; m22 = 1 - m22;
                const double idet = sb_rcp(m11 * m22 - m12 * m21);
                const double idt = sb_rcp(dt);
                double c1p = 0, c2p = 0;
                double* const KC = xs + SB_XS_KC; // k_j with the Woodbury correction applied: every later use is on the thread's own row, so no barrier guards it
                SB_PROF(5)
#if SB_SPLIT_UNROLL
#pragma unroll
#endif
                for (int s = 1; s < 8; s++) {
                    double* ks = K + s * SB_N;
                    const double hd_ = dt * cd[s];
                    double* kprev = ks - SB_N;
                    double h1, h2;
                    sb_hub_dots(S, bs, kprev, h1, h2, lane); // (every warp, from the complete uncorrected k_{s-1}: same values)
                    c1p = (m22 * h1 - m12 * h2) * idet; c2p = (-m21 * h1 + m11 * h2) * idet;
                    const bool inc = s >= 6;
                    const int nold = s - 1;
                    double ua = inc ? U[ic] : u[ic], Racc = 0;
                    // branch-free: all six loads at once, absent terms enter as exact zeros (same bits as skipping them); with a branch per term each
                    // load waited behind the previous term, and the hub butterfly above could not overlap any of it
                    double kjv[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) kjv[j] = KC[j * SB_N + ic];
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        const double kj = (nold > j) ? kjv[j] : 0.0;
                        const double a_ = cA[s][j], c_ = cC[s][j] * idt;
                        if (j < 4) ua = fma(inc ? 0.0 : a_, kj, ua);
                        Racc = fma(c_, kj, Racc);
                    }
                    {
                        const double al = cA[s][s - 1], cl = cC[s][s - 1] * idt;
                        const double kp_ = kprev[ic] + Zp[ic] * c1p + Zq[ic] * c2p;
                        ua = fma(al, kp_, ua);
                        Racc = fma(cl, kp_, Racc);
                        if (has) { KC[(s - 1) * SB_N + i0] = kp_; U[i0] = ua; }
                    }
                    sb_rows_sync();
                    SB_PROF(6)
                    sb_eval_f<true, SPL>(S, bs + cslot[s] * SB_NB, U, ks, lane, &Racc, hd_, dT, nullptr, rb); nf++;
                    sb_rows_sync();
                    SB_PROF(7)
                    if (w0) { double* const k1[1] = {ks}; sb_bsolve<1>(S, k1, di, up, mm, blk, lane); }
                    nsolve++;
                    sb_rows_sync();
                    SB_PROF(8)
                }
                if (w0) for (int i = lane; i < SB_TOPSTORE; i += SB_WARP) blk[i] = 0; // the top block's storage is clean for the next factorisation (its last reader, the solve above, is warp 0 itself)
                {
                    double s1, s2;
                    sb_hub_dots(S, bs, K + 7 * SB_N, s1, s2, lane);
                    c1p = (m22 * s1 - m12 * s2) * idet; c2p = (-m21 * s1 + m11 * s2) * idet;
                }
                const double k8 = sb_wcorr(K[7 * SB_N + ic], Zp[ic], c1p, Zq[ic], c2p);
                // error estimate (own row; the sum over rows in the single-warp order)
                double es = 0; bool bad = false;
                if (has) {
                    const double un = U[i0] + k8;
#if SB_NORMRCP
                    es = k8 * sb_rcp(abstol + reltol * fmax(fabs(u[i0]), fabs(un)));
#else
                    es = k8 / (abstol + reltol * fmax(fabs(u[i0]), fabs(un)));
#endif
                }
                const double EEst = sb_rms(sb_split_sumsq(es, xs, rb, lane));
                SB_PROF(9)
                if (!isfinite(EEst)) { nreject++; dt /= 5; if (has) mm[i0] = 0; if (dt < 1e-15 * fabs(t)) { rc = SB_RC_UNSTABLE; break; } continue; } // (0·NaN may sit in structurally-zero slots)
                if (EEst > 1) { // rejected: nothing to overlap, the answer is needed at once
                    nreject++;
                    ctl_send(EEst, lqold);
                    dt = ctl_resolve(EEst, true, dt);
                    if (dt < 1e-15 * fabs(t)) { rc = SB_RC_DTMIN; break; }
                    continue;
                }
                naccept++;
                if (!last) { ctl_send(EEst, lqold); pend = true; Epend = EEst; } // the new step size is picked up after f0, dT of the next attempt
                const double tn = last ? tend : t + dt;
                const double un_ = has ? U[i0] + k8 : 0.0;
                if (isave < A.nsave && CC.saveat[isave] <= tn) { // dense output (4th order)
                    double a1 = 0, a2 = 0, a3 = 0;
                    for (int j = 0; j < 8; j++) { const double kj = (j < 7) ? KC[j * SB_N + ic] : k8; a1 = fma(cH[0][j], kj, a1); a2 = fma(cH[1][j], kj, a2); a3 = fma(cH[2][j], kj, a3); }
                    while (isave < A.nsave && CC.saveat[isave] <= tn) {
                        const double ts = CC.saveat[isave];
                        const double th = (ts - t) / dt, t1 = 1 - th;
                        const double v = (ts == tn) ? un_ : sb_dense(th, t1, u[ic], un_, a1, a2, a3);
                        if (has) { if (usave) usave[(size_t)isave * SB_N + sb_nat[i0]] = v; if (SRC) di[i0] = v; }
                        sb_rows_sync();
                        emit(di, ts, true);
                    }
                }
                if (has) { if (isnan(un_)) bad = true; }
                bad = sb_rows_or(bad);
                if (has) u[i0] = un_;
                t = tn;
                if (bad) { rc = SB_RC_UNSTABLE; break; }
                if (last) break;
                // the basis at the new time is the one of the last stage time (c = 1): slot 5 -> slot 0, with the derivative kept by the sweep
                for (int m = tid; m < SB_NB; m += NT) { bs[m] = bs[5 * SB_NB + m]; bdv[m] = xs[SB_XS_BD5 + m]; }
                jt = jend;
                moved = true;
                SB_PROF(11)
            }
            if (pend) { asm volatile("bar.sync 4, %0;" ::"n"(SB_SPLIT_THREADS) : "memory"); pend = false; } // (left the loop with an answer outstanding: take it off the barrier)
        }
        sb_rows_sync();
        if (has) A.uend[(size_t)mode * SB_N + sb_nat[i0]] = u[i0];
        for (; isave < A.nsave;) { // save times the mode never reached (failed solve)
            if (usave && has) usave[(size_t)isave * SB_N + i0] = NAN;
            emit(u, t, false);
        }
        SB_PROF_OUT
        if (tid == 0) { A.retcode[mode] = rc; A.stats[4 * mode] = naccept; A.stats[4 * mode + 1] = nreject; A.stats[4 * mode + 2] = nf; A.stats[4 * mode + 3] = nsolve; }
    }
}
#else
#define SB_HAVE_SPLIT 0
#endif

// ================================================================================================ TRBDF2
// The reference's `ptalg(prob; accuracy = 0)` (src/solve.jl:333-335) integrates the perturbations with OrdinaryDiffEq.jl's TRBDF2 instead of Rodas5P.
// OrdinaryDiffEq is an un-vendored dependency and there is no Julia here, so this is the PUBLISHED scheme (Bank et al. 1985, Hosea & Shampine 1996)
// on this engine's linear algebra -- parity with Julia's step selection is unpinned; the oracle carries the same restatement (oracle/sbref.cpp, TRBDF2).
// The system is linear in u: each implicit stage is ONE linear solve with J at the stage time (a converged Newton iteration), W = I/(d dt) − J = B − pφᵀ − qψᵀ.
//   γ = 2 − √2, d = γ/2, ω = √2/4, x = 1/(d dt)
//   stage 2 (trapezoidal rule to t + γ dt):  (x − J_γ) u_γ = x u_n + f_n,            f_γ = x (u_γ − u_n) − f_n
//   stage 3 (BDF2 to t + dt):                (x − J_1) u_1 = x r,  r = a u_γ − b u_n,  f_1 = x (u_1 − r)   (first-same-as-last: no f-evaluation per step)
//   error estimate: (x − J_1)^{-1} x dt Σ (b_i − b̂_i) f_i, third-order companion b̂ = ((1−ω)/3, (3ω+1)/3, d/3); Gustafsson's predictive controller
//   dense output / save points: cubic Hermite on (u_n, f_n, u_1, f_1); sources at the save times as in sb_integrate_kernel.
// One warp per mode, atomic queue, same shared-memory layout as sb_integrate_kernel (K holds f_n, f_γ, f_1, r, the estimate and u_1).
//
// KenCarp4 (ALG = 2; `ptalg(prob; accuracy = 1)`, src/solve.jl:336-337, the algorithm of the reference's sparse-Jacobian test, test/runtests.jl:580-590): the
// ESDIRK half of Kennedy & Carpenter's ARK4(3)6L[2]SA (the reference integrates the whole right-hand side implicitly), γ = 1/4, six stages, stiffly accurate:
//   stage i = 2..6:  (x − J_i) U_i = x (u_n + dt Σ_{j<i} a_ij f_j),  f_i = x (U_i − rhs_i),  x = 1/(γ dt);  u_1 = U_6, f_1 = f_6 (first-same-as-last)
//   error estimate (x − J_6)^{-1} x dt Σ (b_j − b̂_j) f_j (O(dt⁴)); the same controller and dense output.  K holds f_1..f_6, the stage right-hand side and the estimate.
// Same caveat: published tableau (row sums checked in tests/test_host_cpu.py), parity with OrdinaryDiffEq.jl's step selection unpinned.
__constant__ double cKA[6][5] = {{0, 0, 0, 0, 0},
                                 {0.25, 0, 0, 0, 0},
                                 {8611.0 / 62500, -1743.0 / 31250, 0, 0, 0},
                                 {5012029.0 / 34652500, -654441.0 / 2922500, 174375.0 / 388108, 0, 0},
                                 {15267082809.0 / 155376265600.0, -71443401.0 / 120774400, 730878875.0 / 902184768, 2285395.0 / 8070912, 0},
                                 {82889.0 / 524892, 0, 15625.0 / 83664, 69875.0 / 102672, -2260.0 / 8211}};
__constant__ double cKc[6] = {0, 0.5, 83.0 / 250, 31.0 / 50, 17.0 / 20, 1};
__constant__ double cKe[6] = {82889.0 / 524892 - 4586570599.0 / 29645900160.0, 0, 15625.0 / 83664 - 178811875.0 / 945068544, 69875.0 / 102672 - 814220225.0 / 1159782912,
                              -2260.0 / 8211 + 3700637.0 / 11593932, 0.25 - 61727.0 / 225920}; // b − b̂
template <int ALG>
__global__ void __launch_bounds__(SB_WARP, 4) sb_sdirk_kernel(const __grid_constant__ SbSolveArgs A) {
    constexpr int ORDER = (ALG == 1) ? 2 : 4;
    constexpr double EXPO = 1.0 / ((ALG == 1) ? 3 : 4); // exponent of the O(dt³) / O(dt⁴) error estimate in the controller
    extern __shared__ __align__(16) double sm_all[];
    const int lane = threadIdx.x & 31;
    double* const sm = sm_all;
    double *u = sm + SB_SM_U, *U = sm + SB_SM_UU, *K = sm + SB_SM_K, *di = sm + SB_SM_DI, *up = sm + SB_SM_UP, *mm = sm + SB_SM_MM, *Zp = sm + SB_SM_ZP, *Zq = sm + SB_SM_ZQ,
           *bs = sm + SB_SM_BS, *blk = sm + SB_SM_BLK, *kp = sm + SB_SM_KP;
    double* const sbuf = sm + SB_SM_SBUF;
    double *f0 = K, *fg = K + SB_N, *f1 = K + (ALG == 1 ? 2 : 5) * SB_N, *rv = K + (ALG == 1 ? 3 : 6) * SB_N, *ev = K + (ALG == 1 ? 4 : 7) * SB_N, *un = (ALG == 1) ? K + 5 * SB_N : U;
    const double reltol = A.reltol, abstol = A.abstol;
    const bool SRC = A.S != nullptr;
    SbLane S;
    S.load(lane);
    const SbCosmo& CC = A.c0;
    const double G = 2.0 - 1.4142135623730951, D = 0.5 * G, OM = 0.25 * 1.4142135623730951;
    const double BA = 1.0 / (G * (2 - G)), BB = (1 - G) * (1 - G) / (G * (2 - G));
    const double E1 = (4 * OM - 1) / 3, E2 = -1.0 / 3, E3 = 2 * D / 3;
    while (true) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(A.queue, 1);
        qi = __shfl_sync(SB_FULL, qi, 0);
        if (qi >= A.nk) break;
        const int mode = A.order ? A.order[qi] : qi;
        const double k = A.ks[mode];
        double t = A.tini[mode];
        const double tend = CC.tend;
        long long naccept = 0, nreject = 0, nf = 0, nsolve = 0;
        int rc = SB_RC_SUCCESS, isave = 0, wstart = 0;
        double* usave = A.usave ? A.usave + (size_t)mode * A.nsave * SB_N : nullptr;
        double* Sout = SRC ? A.S + (size_t)mode * A.nS * A.nsave : nullptr;
        __syncwarp();
        if (!(k > 0) || !isfinite(k)) {
            for (int i = lane; i < SB_N; i += SB_WARP) A.uend[(size_t)mode * SB_N + i] = NAN;
            if (usave) for (int i = lane; i < A.nsave * SB_N; i += SB_WARP) usave[i] = NAN;
            if (SRC) for (int i = lane; i < A.nsave * A.nS; i += SB_WARP) Sout[i] = NAN;
            if (lane == 0) { A.retcode[mode] = SB_RC_UNSTABLE; for (int j = 0; j < 4; j++) A.stats[4 * mode + j] = 0; }
            continue;
        }
        if (lane < 7) kp[lane] = pow(k, (double)(lane - 3));
        if (lane == 0) sb_initial_state(CC.spl, CC.P, t, k, U);
        __syncwarp();
        for (int i = lane; i < SB_N; i += SB_WARP) u[i] = U[sb_nat[i]];
        __syncwarp();
        // one save point from the state `st` (shared memory, integrator order)
        auto emit = [&](const double* st, double ts, bool valid) {
            if (SRC) {
                double* so = sbuf + (isave & (SB_SWIN - 1));
                if (valid) sb_source_point(S, CC.srcbg + (size_t)isave * SB_SRCBG_STRIDE, kp, k, ts, CC.taurec, A.scale_k, A.nS, st, up, bs, bs + SB_NB, so, SB_SWIN, lane);
                else if (lane < A.nS) so[lane * SB_SWIN] = NAN;
                sb_source_flush(sbuf, Sout, A.nS, A.nsave, isave, wstart, false, lane);
            }
            isave++;
        };
        while (isave < A.nsave && CC.saveat[isave] <= t) {
            const bool at = CC.saveat[isave] == t;
            if (usave) for (int i = lane; i < SB_N; i += SB_WARP) usave[(size_t)isave * SB_N + sb_nat[i]] = at ? u[i] : NAN;
            emit(u, t, at);
        }
        if (tend > t) {
            int jt = sb_interval(CC.tb, t);
            jt = sb_basis_at(S, CC.tb, t, jt, kp, bs, nullptr, lane);
            __syncwarp();
            sb_eval_f<false>(S, bs, u, f0, lane); nf++;
            double dt;
            { // automatic initial step (Hairer) with the method's order
                double d0 = 0, d1 = 0;
                for (int i = lane; i < SB_N; i += SB_WARP) { const double sk = abstol + fabs(u[i]) * reltol; const double a0 = u[i] / sk, a1 = f0[i] / sk; d0 = fma(a0, a0, d0); d1 = fma(a1, a1, d1); }
                d0 = sqrt(warp_sum(d0) / SB_N); d1 = sqrt(warp_sum(d1) / SB_N);
                const double dtmax = tend - t;
                double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                dt0 = fmin(dt0, dtmax);
                for (int i = lane; i < SB_N; i += SB_WARP) U[i] = u[i] + dt0 * f0[i];
                sb_basis_at(S, CC.tb, t + dt0, jt, kp, bs + SB_NB, nullptr, lane);
                __syncwarp();
                sb_eval_f<false>(S, bs + SB_NB, U, fg, lane); nf++;
                double d2 = 0;
                for (int i = lane; i < SB_N; i += SB_WARP) { const double sk = abstol + fabs(u[i]) * reltol; const double r = (fg[i] - f0[i]) / sk; d2 = fma(r, r, d2); }
                d2 = sqrt(warp_sum(d2) / SB_N) / dt0;
                const double dm = fmax(d1, d2);
                const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2 + log10(dm)) / (ORDER + 1.0));
                dt = fmin(fmin(100 * dt0, dt1), dtmax);
                __syncwarp();
            }
            double qold = 1, dtacc = 0, erracc = 0; long long success_iter = 0; // Gustafsson's predictive controller (oracle: PredictiveController)
            double m11 = 0, m12 = 0, m21 = 0, m22 = 0, idet = 0;
            // factor x − J(ts) with the basis of time ts left in slot 0, and the Woodbury data (Z = B^{-1}[p q], 2×2 inverse); solves r in place
            auto factor_solve = [&](double ts, double x, int jh, double* r) {
                const int j = sb_basis_at(S, CC.tb, ts, jh, kp, bs, nullptr, lane);
                __syncwarp();
                sb_factor(S, x, bs, di, up, mm, blk, lane);
#pragma unroll
                for (int rr = 0; rr < SB_R; rr++) {
                    const int i = rr * 32 + lane;
                    if (i < SB_N) { Zp[i] = S.pqc[2 * rr] * bs[SB_LO16(S.pqi[rr])]; Zq[i] = S.pqc[2 * rr + 1] * bs[SB_HI16(S.pqi[rr])]; }
                }
                __syncwarp();
                { double* const zz[3] = {Zp, Zq, r}; sb_bsolve<3>(S, zz, di, up, mm, blk, lane); }
                nsolve += 3;
                sb_hub_dots(S, bs, Zp, m11, m21, lane);
                sb_hub_dots(S, bs, Zq, m12, m22, lane);
                m11 = 1 - m11; m12 = -m12; m21 = -m21; m22 = 1 - m22;
                idet = sb_rcp(m11 * m22 - m12 * m21);
                return j;
            };
            auto woodbury = [&](double* r) { // r += Z M^{-1} [φᵀr; ψᵀr]
                double h1, h2;
                sb_hub_dots(S, bs, r, h1, h2, lane);
                const double c1 = (m22 * h1 - m12 * h2) * idet, c2 = (-m21 * h1 + m11 * h2) * idet;
                __syncwarp();
                for (int i = lane; i < SB_N; i += SB_WARP) r[i] = sb_wcorr(r[i], Zp[i], c1, Zq[i], c2);
                __syncwarp();
            };
            for (int it = 0;; it++) {
                if (it >= A.maxiters) { rc = SB_RC_MAXITERS; break; }
                bool last = false;
                if (t + dt >= tend - 100 * 2.2e-16 * fabs(tend)) { dt = tend - t; last = true; }
                const double x = 1.0 / ((ALG == 1 ? D : 0.25) * dt);
                int jn = jt;
                if (ALG == 1) {
                    for (int i = lane; i < SB_N; i += SB_WARP) U[i] = fma(x, u[i], f0[i]);
                    __syncwarp();
                    const int jg = factor_solve(t + G * dt, x, jt, U);
                    woodbury(U);
                    for (int i = lane; i < SB_N; i += SB_WARP) { fg[i] = x * (U[i] - u[i]) - f0[i]; const double r = BA * U[i] - BB * u[i]; rv[i] = r; un[i] = x * r; }
                    __syncwarp();
                    jn = factor_solve(last ? tend : t + dt, x, jg, un);
                    woodbury(un);
                    for (int i = lane; i < SB_N; i += SB_WARP) { const double f = x * (un[i] - rv[i]); f1[i] = f; ev[i] = x * (dt * (E1 * f0[i] + E2 * fg[i] + E3 * f)); }
                    __syncwarp();
                } else {
#pragma unroll 1
                    for (int sg = 1; sg < 6; sg++) { // stages 2..6; f_1 = f0 sits in K[0..N)
                        for (int i = lane; i < SB_N; i += SB_WARP) {
                            double v = 0;
                            for (int j = 0; j < sg; j++) v = fma(cKA[sg][j], K[j * SB_N + i], v);
                            const double r = fma(dt, v, u[i]);
                            rv[i] = r; U[i] = x * r;
                        }
                        __syncwarp();
                        jn = factor_solve((sg == 5) ? (last ? tend : t + dt) : t + cKc[sg] * dt, x, jn, U);
                        woodbury(U);
                        for (int i = lane; i < SB_N; i += SB_WARP) K[sg * SB_N + i] = x * (U[i] - rv[i]);
                        __syncwarp();
                    }
                    for (int i = lane; i < SB_N; i += SB_WARP) {
                        double v = 0;
                        for (int j = 0; j < 6; j++) v = fma(cKe[j], K[j * SB_N + i], v);
                        ev[i] = x * (dt * v);
                    }
                    __syncwarp();
                }
                { double* const e1[1] = {ev}; sb_bsolve<1>(S, e1, di, up, mm, blk, lane); } nsolve++;
                woodbury(ev);
                double es = 0; bool bad = false;
                for (int i = lane; i < SB_N; i += SB_WARP) { const double r = ev[i] * sb_rcp(abstol + reltol * fmax(fabs(u[i]), fabs(un[i]))); es = fma(r, r, es); }
                const double EEst = sqrt(warp_sum(es) / SB_N);
                if (!isfinite(EEst)) { nreject++; dt /= 5; if (dt < 1e-15 * fabs(t)) { rc = SB_RC_UNSTABLE; break; } continue; }
                const double q = qold = (EEst == 0.0) ? 0.1 : fmax(0.1, fmin(5.0, pow(EEst, EXPO) / 0.9));
                if (EEst > 1) { nreject++; dt = (success_iter == 0) ? 0.1 * dt : dt / qold; if (dt < 1e-15 * fabs(t)) { rc = SB_RC_DTMIN; break; } continue; }
                naccept++;
                double qacc = q;
                if (success_iter > 0) {
                    double qgus = (dtacc / dt) * pow(EEst * EEst / erracc, EXPO);
                    qgus = fmax(0.1, fmin(5.0, qgus / 0.9));
                    qacc = fmax(q, qgus);
                }
                if (1.0 <= qacc && qacc <= 1.2) qacc = 1;
                success_iter++;
                dtacc = dt; erracc = fmax(1e-2, EEst);
                const double dtnew = dt / qacc;
                const double tn = last ? tend : t + dt;
                while (isave < A.nsave && CC.saveat[isave] <= tn) { // cubic Hermite on (u_n, f_n, u_1, f_1)
                    const double ts = CC.saveat[isave];
                    const double th = (ts - t) / dt;
                    const bool atend = ts == tn;
                    for (int i = lane; i < SB_N; i += SB_WARP) {
                        const double du = un[i] - u[i];
                        const double v = atend ? un[i] : (1 - th) * u[i] + th * un[i] + th * (th - 1) * ((1 - 2 * th) * du + (th - 1) * dt * f0[i] + th * dt * f1[i]);
                        if (usave) usave[(size_t)isave * SB_N + sb_nat[i]] = v;
                        if (SRC) di[i] = v;
                    }
                    __syncwarp();
                    emit(di, ts, true);
                }
                for (int i = lane; i < SB_N; i += SB_WARP) { const double v = un[i]; if (isnan(v)) bad = true; u[i] = v; f0[i] = f1[i]; }
                t = tn;
                bad = __any_sync(SB_FULL, bad);
                __syncwarp();
                if (bad) { rc = SB_RC_UNSTABLE; break; }
                if (last) break;
                dt = dtnew;
                jt = jn;
            }
        }
        __syncwarp();
        for (int i = lane; i < SB_N; i += SB_WARP) A.uend[(size_t)mode * SB_N + sb_nat[i]] = u[i];
        while (isave < A.nsave) { // save times the mode never reached (failed solve)
            if (usave) for (int i = lane; i < SB_N; i += SB_WARP) usave[(size_t)isave * SB_N + i] = NAN;
            emit(u, t, false);
        }
        if (lane == 0) { A.retcode[mode] = rc; A.stats[4 * mode] = naccept; A.stats[4 * mode + 1] = nreject; A.stats[4 * mode + 2] = nf; A.stats[4 * mode + 3] = nsolve; }
    }
}

// Δm(τ,k) for P(k): one thread per mode
__global__ void sb_deltam_kernel(const double* __restrict__ P, SbSpline spl, double tau, int nk, const double* __restrict__ ks, const double* __restrict__ u, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nk) return;
    double y[5]; sb_spl_eval(spl, tau, y, nullptr);
    out[i] = sb_delta_m(tau, ks[i], y, P, u + (size_t)i * SB_N);
}

// Per-τ background quantities for the CMB sources: one thread per saved time (row layout: see SB_SRCBG_STRIDE above).
__global__ void sb_srcbg_kernel(const double* __restrict__ P, SbSpline spl, int nt, const double* __restrict__ taus, double* __restrict__ srcbg) {
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= nt) return;
    double tau = taus[it], y[5], g[5], kd[3];
    sb_spl_eval(spl, tau, y, nullptr);
    sb_kappa_derivs(y, P, g, kd);
    double* o = srcbg + (size_t)it * SB_SRCBG_STRIDE;
    o[0] = kd[0]; o[1] = kd[1]; o[2] = kd[2]; o[3] = exp(-(y[1] - P[SB_P_KAPPA0])); o[4] = P[SB_P_TAU0] - tau; o[5] = 0; o[6] = 0; o[7] = 0;
    sb_beta(tau, y, g, P, o + 8, o + 8 + SB_NBETA); // time derivatives along the background flow: dy/dτ = g(y)
}

// CMB source functions at every saved (k, τ) from states kept in HBM: one thread per point (the fused path inside the integrator,
// sb_source_point, is what the C_l pipeline uses; this kernel serves callers that asked for the states themselves).
// out layout: S[ik][iS][it], iS = 0: ST, 1: SE, (nS == 3) 2: Sψ
__global__ void sb_source_kernel(int nt, const double* __restrict__ taus, const double* __restrict__ srcbg, int nk, const double* __restrict__ ks,
                                 const double* __restrict__ usave, double* __restrict__ S, int scale_k, int nS, double taurec) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nk * nt) return;
    int ik = idx / nt, it = idx % nt;
    const double k = ks[ik];
    const double* sb = srcbg + (size_t)it * SB_SRCBG_STRIDE;
    const double* u = usave + ((size_t)ik * nt + it) * SB_N;
    double kpw[7];
    for (int e = 0; e < 7; e++) kpw[e] = pow(k, (double)(e - 3));
    double b[SB_NB], bd[SB_NB];
    for (int m = 0; m < SB_NB; m++) { double kk = kpw[sb_basis_kpow[m] + 3]; b[m] = kk * sb[8 + sb_basis_beta[m]]; bd[m] = kk * sb[8 + SB_NBETA + sb_basis_beta[m]]; }
    const double Phd = sb_hub_seq(2, b, u), Psi = sb_hub_seq(3, b, u);
    double ud[SB_N];
    for (int i = 0; i < SB_N; i++) ud[i] = sb_row_dot(i, b, u);
    sb_source_tail(sb, k, taus[it], taurec, scale_k, nS, b, bd, u, ud, Phd, Psi, S + (size_t)ik * nS * nt + it, nt);
}

// ================================================================================================ C ABI (device pointers)
#define SB_CUDA_CHECK(x)                                             \
    do {                                                             \
        cudaError_t e_ = (x);                                        \
        if (e_ != cudaSuccess) { fprintf(stderr, "symboltz_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -(int)e_ - 1000; } \
    } while (0)

extern "C" {

int sbm_smem_bytes(void) { return SB_SM_BYTES; }
int sbm_srcbg_stride(void) { return SB_SRCBG_STRIDE; }

// dtab: ((nb-1)·msub + 1)·2·NBETA doubles
// Batched background solve on the device: dP [n][NPAR] (kappa0/tau0 slots are filled in), outputs dt [n][cap], dy/ddy [n][cap][5],
// dinfo [n][8] (tau0, kappa0, taurec, retcode, naccept, nreject, 0, 0), dnb [n] (knots, or -1 if cap was too small).
int sbm_solvebg_batch(int n, double* dP, double tini, double tmax, double reltol, double abstol, int cap, double* dt, double* dy, double* ddy, double* dinfo, int* dnb, void* stream) {
    if (n <= 0) return 0;
    if (cap < 2) return -1;
    SB_CUDA_CHECK(cudaFuncSetAttribute(sb_solvebg_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 0)); // local-memory working set: all of L1
    sb_solvebg_kernel<<<(n + 31) / 32, 32, 0, (cudaStream_t)stream>>>(n, dP, tini, tmax, reltol, abstol, cap, dt, dy, ddy, dinfo, dnb);
    SB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sbm_build_table(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, double* dtab, void* stream) {
    SbSpline spl{nb, dt, dy, ddy};
    int nnode = (nb - 1) * msub + 1;
    sb_table_kernel<<<(nnode + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dP, spl, msub, dtab);
    SB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Perturbation solve over nk modes (reference solvept, src/solve.jl:543-569).  All array arguments are DEVICE pointers.
// dorder may be NULL (natural order).  dqueue: one int, zeroed by this call.  nctas <= 0: fill the GPU.
static int sb_solvept_impl(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                           const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat, double reltol, double abstol,
                           int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, int nctas, void* stream, double* dtrace, int ntrace,
                           const int* ditems, const int* dibeg, int nlists, double* dcont, int* dflags, const SbCosmo* dcosmos = nullptr, const int* dcosmo_of = nullptr,
                           const sbm_src_t* src = nullptr) {
    if (nk <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SbSolveArgs A;
    const bool fused = src && src->dS && nsave > 0;
    if (fused && (src->nS < 2 || src->nS > 3 || (!dcosmos && !src->dsrcbg))) return -1;
    A.c0 = SbCosmo{dP, SbSpline{nb, dt, dy, ddy}, SbTable{nb, msub, nlut, s0, 1.0 / dsl, dt, dlut, dtab}, tend, dsaveat, fused ? src->dsrcbg : nullptr, fused ? src->taurec : 0.0};
    A.cosmos = dcosmos; A.cosmo_of = dcosmo_of;
    A.S = fused ? src->dS : nullptr; A.nS = fused ? src->nS : 0; A.scale_k = fused ? src->scale_k : 0;
    A.nk = nk; A.ks = dks; A.tini = dtini; A.order = dorder; A.nsave = (dusave || fused) ? nsave : 0;
    A.reltol = reltol; A.abstol = abstol; A.maxiters = maxiters; A.usave = dusave; A.uend = duend; A.retcode = dretcode; A.stats = dstats; A.queue = dqueue; A.trace = dtrace; A.ntrace = ntrace;
    A.items = ditems; A.ibeg = dibeg; A.cont = dcont; A.flags = dflags;
    SB_CUDA_CHECK(cudaMemsetAsync(dqueue, 0, sizeof(int), st));
    static int occ = 0, occb = 0, nsm = 0;
    if (!occ) {
        int dev; SB_CUDA_CHECK(cudaGetDevice(&dev));
        SB_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        SB_CUDA_CHECK(cudaFuncSetAttribute(sb_integrate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SM_BYTES));
        SB_CUDA_CHECK(cudaFuncSetAttribute(sb_integrate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SM_BYTES));
        int occ_b = 0;
        SB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sb_integrate_kernel<false>, SB_WARP * SB_WARPS_PER_CTA, SB_SM_BYTES));
        SB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, sb_integrate_kernel<true>, SB_WARP * SB_WARPS_PER_CTA, SB_SM_BYTES));
        occ = std::max(1, occ); occb = std::max(1, occ_b);
    }
    int grid = nctas > 0 ? nctas : std::min((nk + SB_WARPS_PER_CTA - 1) / SB_WARPS_PER_CTA, nsm * (dcosmos ? occb : occ));
    if (ditems) {
        // static schedule: one list per warp; a continuation item spins on a flag set by the warp that runs the first piece,
        // which is the FIRST item of its list -- so every list must be resident from the start
        if (nlists <= 0 || nlists > nsm * (dcosmos ? occb : occ) * SB_WARPS_PER_CTA || nlists % SB_WARPS_PER_CTA) return -3;
        grid = nlists / SB_WARPS_PER_CTA;
        SB_CUDA_CHECK(cudaMemsetAsync(dflags, 0, sizeof(int) * nk, st));
    }
    if (dcosmos) sb_integrate_kernel<true><<<grid, SB_WARP * SB_WARPS_PER_CTA, SB_SM_BYTES, st>>>(A);
    else sb_integrate_kernel<false><<<grid, SB_WARP * SB_WARPS_PER_CTA, SB_SM_BYTES, st>>>(A);
    SB_CUDA_CHECK(cudaGetLastError());
    return grid;
}

int sbm_solvept(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat, double reltol, double abstol,
                int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, int nctas, void* stream, double* dtrace, int ntrace) {
    return sb_solvept_impl(dP, nb, dt, dy, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, dorder, tend, nsave, dsaveat, reltol, abstol, maxiters, dusave, duend, dretcode, dstats, dqueue, nctas,
                           stream, dtrace, ntrace, nullptr, nullptr, 0, nullptr, nullptr);
}

// Same solve under a static schedule built by the host from a cost model (McNaughton's wrap-around rule: the modes are laid end to
// end in units of estimated attempts and cut into `nlists` equal chunks, one per resident warp; a mode that straddles a cut runs
// its first attempts as the FIRST item of the next warp's list and is finished as the LAST item of the previous one).
// ditems[nitems][3] = (mode, quota, cont); dibeg[nlists + 1]; dcont: nk * sbm_cont_stride() doubles; dflags: nk ints.
int sbm_solvept_sched(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                      const double* dks, const double* dtini, double tend, int nsave, const double* dsaveat, double reltol, double abstol, int maxiters, double* dusave, double* duend,
                      int* dretcode, long long* dstats, int* dqueue, const int* ditems, const int* dibeg, int nlists, double* dcont, int* dflags, void* stream) {
    if (!ditems || !dibeg || !dcont || !dflags) return -1;
    return sb_solvept_impl(dP, nb, dt, dy, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, nullptr, tend, nsave, dsaveat, reltol, abstol, maxiters, dusave, duend, dretcode, dstats, dqueue, 0,
                           stream, nullptr, 0, ditems, dibeg, nlists, dcont, dflags);
}
// One launch over the (cosmology, mode) pairs of a parameter sweep (SURVEY §8b "batched variants with leading ncosmo dimension";
// the reference loops `spectrum_matter(probgen(θ), ks)` serially, docs/src/forecasting.md:56-59).  dcosmos: device array of ncosmo
// sbm_cosmo_t; dcosmo_of[nk]: cosmology of each mode; dorder: queue order (NULL = natural), or a static schedule (ditems != NULL).
// dusave (optional): [nk][nsave][N] with every cosmology supplying its own nsave save times.
int sbm_solvept_batch(int ncosmo, const void* dcosmos, int nk, const double* dks, const double* dtini, const int* dcosmo_of, const int* dorder, int nsave, double reltol, double abstol,
                      int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const int* ditems, const int* dibeg, int nlists, double* dcont, int* dflags,
                      void* stream) {
    if (ncosmo <= 0 || !dcosmos || !dcosmo_of) return -1;
    if (ditems && (!dibeg || !dcont || !dflags)) return -1;
    return sb_solvept_impl(nullptr, 0, nullptr, nullptr, nullptr, 1, 1, 0.0, 1.0, nullptr, nullptr, nk, dks, dtini, dorder, 0.0, nsave, nullptr, reltol, abstol, maxiters, dusave, duend, dretcode, dstats,
                           dqueue, 0, stream, nullptr, 0, ditems, dibeg, nlists, dcont, dflags, (const SbCosmo*)dcosmos, dcosmo_of);
}
// The three solves with the CMB source functions formed INSIDE the integrator at the save times (reference: the output_func of
// solvept evaluates getsym(prob.pt, Ss) on the dense output and keeps only S, src/observables/fourier.jl:267-281): `src` (host struct)
// names the output dS[nk][nS][nsave] and, for the single-cosmology calls, the per-save-time background table from sbm_srcbg and
// taurec (the batched call reads both from each sbm_cosmo_t).  dusave may be NULL: the states then never leave the SM
// (the round-1 path wrote usave[nk][nsave][N] = 397 MB per cosmology at the bench size and re-read it in a second kernel).
int sbm_solvept_src(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                    const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat, double reltol, double abstol,
                    int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, int nctas, void* stream, const sbm_src_t* src) {
    return sb_solvept_impl(dP, nb, dt, dy, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, dorder, tend, nsave, dsaveat, reltol, abstol, maxiters, dusave, duend, dretcode, dstats, dqueue, nctas,
                           stream, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, src);
}
int sbm_solvept_sched_src(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                          const double* dks, const double* dtini, double tend, int nsave, const double* dsaveat, double reltol, double abstol, int maxiters, double* dusave, double* duend,
                          int* dretcode, long long* dstats, int* dqueue, const int* ditems, const int* dibeg, int nlists, double* dcont, int* dflags, void* stream, const sbm_src_t* src) {
    if (!ditems || !dibeg || !dcont || !dflags) return -1;
    return sb_solvept_impl(dP, nb, dt, dy, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, nullptr, tend, nsave, dsaveat, reltol, abstol, maxiters, dusave, duend, dretcode, dstats, dqueue, 0,
                           stream, nullptr, 0, ditems, dibeg, nlists, dcont, dflags, nullptr, nullptr, src);
}
int sbm_solvept_batch_src(int ncosmo, const void* dcosmos, int nk, const double* dks, const double* dtini, const int* dcosmo_of, const int* dorder, int nsave, double reltol, double abstol,
                          int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const int* ditems, const int* dibeg, int nlists, double* dcont, int* dflags,
                          void* stream, const sbm_src_t* src) {
    if (ncosmo <= 0 || !dcosmos || !dcosmo_of) return -1;
    if (ditems && (!dibeg || !dcont || !dflags)) return -1;
    return sb_solvept_impl(nullptr, 0, nullptr, nullptr, nullptr, 1, 1, 0.0, 1.0, nullptr, nullptr, nk, dks, dtini, dorder, 0.0, nsave, nullptr, reltol, abstol, maxiters, dusave, duend, dretcode, dstats,
                           dqueue, 0, stream, nullptr, 0, ditems, dibeg, nlists, dcont, dflags, (const SbCosmo*)dcosmos, dcosmo_of, src);
}
// Parameter lanes in lockstep (BASELINE config 5: "ForwardDiff gradient ... via dual-number lanes in the batched solve"; reference:
// ForwardDiff.Dual parameters pushed through solvept, test/runtests.jl:363-422).  G = ncosmo cosmologies -- lane 0 the primal, lane
// j >= 1 with one parameter moved by δ_j -- are integrated for the same nk wavenumbers by CTAs of G warps that share ONE step
// controller whose error norm covers the primal and the partials (u^j − u^0)·invdelta[j] (see sb_integrate_kernel<.., GROUP>).
// Arrays are laid out [mode][lane]: dks/dtini/dcosmo_of have nk·G entries with dks[m·G + j] = k_m and dcosmo_of[m·G + j] = j;
// outputs likewise (duend[nk·G][N], dretcode[nk·G], dstats[nk·G][4], src->dS[nk·G][nS][nsave]).  dorder: optional order of the
// nk groups.  invdelta: host array of G doubles ([0] unused); tend_common: the smallest end time (sbm_cosmo_t.tend) of the lanes -- the
// lockstep phase ends there and every lane closes with one private step to its own end time.  Cost: G solves running side by side.
int sbm_solvept_lanes(int ncosmo, const void* dcosmos, int nk, const double* dks, const double* dtini, const int* dcosmo_of, const int* dorder, int nsave, double reltol, double abstol,
                      int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const double* invdelta, double tend_common, void* stream,
                      const sbm_src_t* src) {
    const int G = ncosmo;
    if (G < 2 || G > SB_GROUP_MAX || !dcosmos || !dcosmo_of || !invdelta || !(tend_common > 0)) return -1;
    if (nk <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SbSolveArgs A;
    memset(&A, 0, sizeof(A));
    const bool fused = src && src->dS && nsave > 0;
    if (fused && (src->nS < 2 || src->nS > 3)) return -1;
    A.c0 = SbCosmo{nullptr, SbSpline{0, nullptr, nullptr, nullptr}, SbTable{0, 1, 1, 0.0, 1.0, nullptr, nullptr, nullptr}, 0.0, nullptr, nullptr, 0.0};
    A.cosmos = (const SbCosmo*)dcosmos; A.cosmo_of = dcosmo_of;
    A.S = fused ? src->dS : nullptr; A.nS = fused ? src->nS : 0; A.scale_k = fused ? src->scale_k : 0;
    A.nk = nk * G; A.ks = dks; A.tini = dtini; A.order = dorder; A.nsave = (dusave || fused) ? nsave : 0;
    A.reltol = reltol; A.abstol = abstol; A.maxiters = maxiters; A.usave = dusave; A.uend = duend; A.retcode = dretcode; A.stats = dstats; A.queue = dqueue;
    A.G = G; A.tendc = tend_common;
    for (int j = 0; j < SB_GROUP_MAX; j++) A.invdelta[j] = j < G ? invdelta[j] : 0.0;
    SB_CUDA_CHECK(cudaMemsetAsync(dqueue, 0, sizeof(int), st));
    const size_t smem = ((size_t)G * SB_SM_DOUBLES + SB_XCH_HDR + (size_t)2 * G * 3 * SB_N) * sizeof(double);
    int dev, nsm, occ = 0;
    SB_CUDA_CHECK(cudaGetDevice(&dev));
    SB_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SB_CUDA_CHECK(cudaFuncSetAttribute(sb_integrate_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sb_integrate_kernel<true, true>, SB_WARP * G, smem));
    const int grid = std::min(nk, nsm * std::max(1, occ));
    sb_integrate_kernel<true, true><<<grid, SB_WARP * G, smem, st>>>(A);
    SB_CUDA_CHECK(cudaGetLastError());
    return grid;
}
// The same solve as sbm_solvept_src with ONE CTA of SB_R warps per mode (sb_integrate_split_kernel): for launches that leave most of the GPU
// idle under the warp-per-mode mapping (fewer modes than sbm_split_capacity()), where the time is the slowest mode's sequential attempts.
// Bit-identical results.  Returns the grid size, -5 if this model has no split kernel, or a negative error.
#ifdef SB_TIME_PROF
int sbm_time_prof(unsigned long long* lists, unsigned long long* modes, int reset) {
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    SB_CUDA_CHECK(cudaMemcpyFromSymbol(lists, sb_tp_list, sizeof(sb_tp_list)));
    SB_CUDA_CHECK(cudaMemcpyFromSymbol(modes, sb_tp_mode, sizeof(sb_tp_mode)));
    if (reset) { void* p_; SB_CUDA_CHECK(cudaGetSymbolAddress(&p_, sb_tp_list)); SB_CUDA_CHECK(cudaMemset(p_, 0, sizeof(sb_tp_list))); SB_CUDA_CHECK(cudaGetSymbolAddress(&p_, sb_tp_mode)); SB_CUDA_CHECK(cudaMemset(p_, 0, sizeof(sb_tp_mode))); }
    return 0;
}
#endif
#ifdef SB_SPLIT_PROF
int sbm_split_prof(unsigned long long* out, int reset) {
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    SB_CUDA_CHECK(cudaMemcpyFromSymbol(out, sb_split_prof, sizeof(sb_split_prof)));
    if (reset) { unsigned long long z[24] = {0}; SB_CUDA_CHECK(cudaMemcpyToSymbol(sb_split_prof, z, sizeof(z))); }
    return 0;
}
#endif
int sbm_split_capacity(void) { // modes that run concurrently under the split mapping (0: not available for this model)
#if SB_HAVE_SPLIT
    int dev, nsm, occ = 0;
    const size_t smem = (size_t)(SB_SM_DOUBLES + SB_XS_DOUBLES) * sizeof(double);
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(sb_integrate_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sb_integrate_split_kernel, SB_SPLIT_THREADS, smem) != cudaSuccess) return 0;
    return nsm * occ;
#else
    return 0;
#endif
}
int sbm_solvept_split(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                      const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat, double reltol, double abstol,
                      int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream, const sbm_src_t* src) {
#if SB_HAVE_SPLIT
    if (nk <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SbSolveArgs A;
    memset(&A, 0, sizeof(A));
    const bool fused = src && src->dS && nsave > 0;
    if (fused && (src->nS < 2 || src->nS > 3 || !src->dsrcbg)) return -1;
    A.c0 = SbCosmo{dP, SbSpline{nb, dt, dy, ddy}, SbTable{nb, msub, nlut, s0, 1.0 / dsl, dt, dlut, dtab}, tend, dsaveat, fused ? src->dsrcbg : nullptr, fused ? src->taurec : 0.0};
    A.S = fused ? src->dS : nullptr; A.nS = fused ? src->nS : 0; A.scale_k = fused ? src->scale_k : 0;
    A.nk = nk; A.ks = dks; A.tini = dtini; A.order = dorder; A.nsave = (dusave || fused) ? nsave : 0;
    A.reltol = reltol; A.abstol = abstol; A.maxiters = maxiters; A.usave = dusave; A.uend = duend; A.retcode = dretcode; A.stats = dstats; A.queue = dqueue;
    SB_CUDA_CHECK(cudaMemsetAsync(dqueue, 0, sizeof(int), st));
    const int cap = sbm_split_capacity();
    if (cap <= 0) return -5;
    const size_t smem = (size_t)(SB_SM_DOUBLES + SB_XS_DOUBLES) * sizeof(double);
    const int grid = std::min(nk, cap);
    sb_integrate_split_kernel<<<grid, SB_SPLIT_THREADS, smem, st>>>(A);
    SB_CUDA_CHECK(cudaGetLastError());
    return grid;
#else
    return -5;
#endif
}
// The single-cosmology solve with TRBDF2 instead of Rodas5P (reference ptalg(prob; accuracy = 0), src/solve.jl:333-335): arguments and results as
// sbm_solvept_src (stats[3] counts the linear solves); one warp per mode, atomic queue.  sbm_solvept_sdirk: alg 1 = TRBDF2, 2 = KenCarp4 (accuracy = 1).
int sbm_solvept_sdirk(int alg, const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                       const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat, double reltol, double abstol,
                       int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream, const sbm_src_t* src) {
    if (nk <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SbSolveArgs A;
    memset(&A, 0, sizeof(A));
    const bool fused = src && src->dS && nsave > 0;
    if (fused && (src->nS < 2 || src->nS > 3 || !src->dsrcbg)) return -1;
    A.c0 = SbCosmo{dP, SbSpline{nb, dt, dy, ddy}, SbTable{nb, msub, nlut, s0, 1.0 / dsl, dt, dlut, dtab}, tend, dsaveat, fused ? src->dsrcbg : nullptr, fused ? src->taurec : 0.0};
    A.S = fused ? src->dS : nullptr; A.nS = fused ? src->nS : 0; A.scale_k = fused ? src->scale_k : 0;
    A.nk = nk; A.ks = dks; A.tini = dtini; A.order = dorder; A.nsave = (dusave || fused) ? nsave : 0;
    A.reltol = reltol; A.abstol = abstol; A.maxiters = maxiters; A.usave = dusave; A.uend = duend; A.retcode = dretcode; A.stats = dstats; A.queue = dqueue;
    SB_CUDA_CHECK(cudaMemsetAsync(dqueue, 0, sizeof(int), st));
    int dev, nsm, occ = 0;
    SB_CUDA_CHECK(cudaGetDevice(&dev));
    SB_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    if (alg != 1 && alg != 2) return -1;
    auto kern = (alg == 1) ? sb_sdirk_kernel<1> : sb_sdirk_kernel<2>;
    SB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SM_BYTES));
    SB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SB_WARP, SB_SM_BYTES));
    const int grid = std::min(nk, nsm * std::max(occ, 1));
    kern<<<grid, SB_WARP, SB_SM_BYTES, st>>>(A);
    SB_CUDA_CHECK(cudaGetLastError());
    return grid;
}
int sbm_solvept_trbdf2(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                       const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat, double reltol, double abstol,
                       int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream, const sbm_src_t* src) {
    return sbm_solvept_sdirk(1, dP, nb, dt, dy, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, dorder, tend, nsave, dsaveat, reltol, abstol, maxiters, dusave, duend, dretcode, dstats, dqueue, stream, src);
}
int sbm_solvept_kencarp4(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut, const double* dtab, int nk,
                         const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat, double reltol, double abstol,
                         int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream, const sbm_src_t* src) {
    return sbm_solvept_sdirk(2, dP, nb, dt, dy, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, dorder, tend, nsave, dsaveat, reltol, abstol, maxiters, dusave, duend, dretcode, dstats, dqueue, stream, src);
}
// Per-save-time background table of the source evaluation: dsrcbg[nt][sbm_srcbg_stride()] (κ̇, κ̈, κ⃛, e^{−κ}, τ0 − τ, the β_m and their
// flow derivatives at dtaus[nt]); input of sbm_solvept_src / sbm_cosmo_t.srcbg and of sbm_sources.
int sbm_srcbg(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int nt, const double* dtaus, double* dsrcbg, void* stream) {
    if (nt <= 0) return 0;
    sb_srcbg_kernel<<<(nt + 63) / 64, 64, 0, (cudaStream_t)stream>>>(dP, SbSpline{nb, dt, dy, ddy}, nt, dtaus, dsrcbg);
    SB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int sbm_cosmo_bytes(void) { return (int)sizeof(SbCosmo); }
int sbm_cont_stride(void) { return SB_CONT; }
int sbm_warps_per_cta(void) { return SB_WARPS_PER_CTA; }
// number of integrator warps that can be resident at once on the current device (= the number of lists of a static schedule)
int sbm_resident_warps(void) {
    int dev, nsm, occ;
    SB_CUDA_CHECK(cudaGetDevice(&dev));
    SB_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SB_CUDA_CHECK(cudaFuncSetAttribute(sb_integrate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SM_BYTES));
    SB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sb_integrate_kernel<false>, SB_WARP * SB_WARPS_PER_CTA, SB_SM_BYTES));
    return nsm * occ * SB_WARPS_PER_CTA;
}
int sbm_resident_warps_batch(void) {
    int dev, nsm, occ;
    SB_CUDA_CHECK(cudaGetDevice(&dev));
    SB_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SB_CUDA_CHECK(cudaFuncSetAttribute(sb_integrate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SM_BYTES));
    SB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sb_integrate_kernel<true>, SB_WARP * SB_WARPS_PER_CTA, SB_SM_BYTES));
    return nsm * occ * SB_WARPS_PER_CTA;
}

int sbm_delta_m(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, double tau, int nk, const double* dks, const double* du, double* dout, void* stream) {
    if (nk <= 0) return 0;
    sb_deltam_kernel<<<(nk + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dP, SbSpline{nb, dt, dy, ddy}, tau, nk, dks, du, dout);
    SB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// sources: dsrcbg is scratch of nt·sbm_srcbg_stride() doubles; dS: [nk][nS][nt], nS = 2 (ST, SE) or 3 (+ lensing Sψ, needs taurec)
int sbm_sources(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int nt, const double* dtaus, double* dsrcbg, int nk, const double* dks,
                const double* dusave, double* dS, int scale_k, int nS, double taurec, void* stream) {
    if (nk <= 0 || nt <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    sb_srcbg_kernel<<<(nt + 63) / 64, 64, 0, st>>>(dP, SbSpline{nb, dt, dy, ddy}, nt, dtaus, dsrcbg);
    SB_CUDA_CHECK(cudaGetLastError());
    int n = nk * nt;
    sb_source_kernel<<<(n + 127) / 128, 128, 0, st>>>(nt, dtaus, dsrcbg, nk, dks, dusave, dS, scale_k, nS, taurec);
    SB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
} // extern "C"

// ---------------------------------------------------------------------------------------------- host-buffer entry point
// Everything above takes device pointers (the host layer owns the device memory).  This one-call variant takes HOST pointers
// only, so that a host language without a CUDA allocator (Julia without CUDA.jl, plain C) can bind the path with a single
// ccall: it uploads the background knots, builds the β-table and the interval look-up, orders the work queue by descending
// k, launches the solve (+ Δm and the CMB sources when asked for), and downloads the results.  Same kernels, same defaults
// (msub = 16, 4096-entry look-up) as the Python host layer, hence identical results.
struct SbDevBuf { // frees on scope exit
    std::vector<void*> ptrs;
    ~SbDevBuf() { for (void* q : ptrs) cudaFree(q); }
    template <class T> cudaError_t get(T** out, size_t n, const T* host = nullptr) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(q);
        *out = (T*)q;
        if (host && n) e = cudaMemcpy(q, host, n * sizeof(T), cudaMemcpyHostToDevice);
        return e;
    }
};

extern "C" {
int sbm_solvept_host(const double* P, int nb, const double* t, const double* y, const double* dy, int nk, const double* ks, const double* tini, double tend, int nsave,
                     const double* saveat, double reltol, double abstol, int maxiters, double* usave, double* uend, int* retcode, long long* stats, double* delta_m, int nS,
                     double taurec, int scale_k, double* S) {
    if (nb < 2 || nk < 0 || nsave < 0 || (S && (nS < 2 || nS > 3 || nsave == 0))) return -1;
    if (nk == 0) return 0;
    const int msub = 16, nlut = 4096;
    const double s0 = log(t[0]), dsl = (log(t[nb - 1]) - s0) / nlut;
    std::vector<int> lut(nlut), order(nk);
    for (int q = 0, j = 0; q < nlut; q++) { // knot interval containing exp(s0 + q·dsl)
        const double tq = exp(s0 + dsl * q);
        while (j < nb - 2 && t[j + 1] <= tq) j++;
        lut[q] = j;
    }
    for (int i = 0; i < nk; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ks[a] > ks[b]; });
    SbDevBuf B;
    double *dP, *dt, *dy_, *ddy, *dtab, *dks, *dtini, *dsave, *dus = nullptr, *due, *ddm = nullptr, *dsb = nullptr, *dS = nullptr;
    int *dlut, *dorder, *dret, *dq;
    long long* dst;
    const bool keep = usave && nsave > 0;      // the states are kept only if the caller wants them
    const bool fuse = S && nsave > 0;          // the sources are formed inside the integrator
    SB_CUDA_CHECK(B.get(&dP, SB_NPAR, P)); SB_CUDA_CHECK(B.get(&dt, nb, t)); SB_CUDA_CHECK(B.get(&dy_, (size_t)nb * 5, y)); SB_CUDA_CHECK(B.get(&ddy, (size_t)nb * 5, dy));
    SB_CUDA_CHECK(B.get(&dtab, ((size_t)(nb - 1) * msub + 1) * 2 * SB_NBETA)); SB_CUDA_CHECK(B.get(&dks, nk, ks)); SB_CUDA_CHECK(B.get(&dtini, nk, tini));
    SB_CUDA_CHECK(B.get(&dsave, nsave, saveat)); SB_CUDA_CHECK(B.get(&dlut, nlut, lut.data())); SB_CUDA_CHECK(B.get(&dorder, nk, order.data()));
    SB_CUDA_CHECK(B.get(&due, (size_t)nk * SB_N)); SB_CUDA_CHECK(B.get(&dret, nk)); SB_CUDA_CHECK(B.get(&dst, (size_t)nk * 4)); SB_CUDA_CHECK(B.get(&dq, 1));
    if (keep) SB_CUDA_CHECK(B.get(&dus, (size_t)nk * nsave * SB_N));
    int rc = sbm_build_table(dP, nb, dt, dy_, ddy, msub, dtab, nullptr);
    if (rc < 0) return rc;
    sbm_src_t src{nullptr, nullptr, nS, scale_k, taurec};
    if (fuse) {
        SB_CUDA_CHECK(B.get(&dsb, (size_t)nsave * SB_SRCBG_STRIDE)); SB_CUDA_CHECK(B.get(&dS, (size_t)nk * nS * nsave));
        rc = sbm_srcbg(dP, nb, dt, dy_, ddy, nsave, dsave, dsb, nullptr);
        if (rc < 0) return rc;
        src.dsrcbg = dsb; src.dS = dS;
    }
    rc = sbm_solvept_src(dP, nb, dt, dy_, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, dorder, tend, (keep || fuse) ? nsave : 0, dsave, reltol, abstol, maxiters, dus, due, dret, dst, dq, 0,
                         nullptr, fuse ? &src : nullptr);
    if (rc < 0) return rc;
    if (delta_m) {
        SB_CUDA_CHECK(B.get(&ddm, nk));
        rc = sbm_delta_m(dP, nb, dt, dy_, ddy, tend, nk, dks, due, ddm, nullptr);
        if (rc < 0) return rc;
    }
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    if (uend) SB_CUDA_CHECK(cudaMemcpy(uend, due, (size_t)nk * SB_N * sizeof(double), cudaMemcpyDeviceToHost));
    if (usave && nsave) SB_CUDA_CHECK(cudaMemcpy(usave, dus, (size_t)nk * nsave * SB_N * sizeof(double), cudaMemcpyDeviceToHost));
    if (retcode) SB_CUDA_CHECK(cudaMemcpy(retcode, dret, (size_t)nk * sizeof(int), cudaMemcpyDeviceToHost));
    if (stats) SB_CUDA_CHECK(cudaMemcpy(stats, dst, (size_t)nk * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    if (delta_m) SB_CUDA_CHECK(cudaMemcpy(delta_m, ddm, (size_t)nk * sizeof(double), cudaMemcpyDeviceToHost));
    if (S) SB_CUDA_CHECK(cudaMemcpy(S, dS, (size_t)nk * nS * nsave * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
} // extern "C"
