/* symboltz_b200.h -- C ABI of the B200-native SymBoltz hot path.
 *
 * Three shared libraries export these symbols:
 *   libsbm_<model>.so  (one per model structure lmax/nx/w0wa; built at problem-build time by symboltz.jl_b200/build.py,
 *                       sources: symboltz.jl_b200/csrc/sb_engine.cu + sb_debug.cpp + generated sb_model_gen.h)   -> sbm_*
 *   libsbl.so          (model independent; symboltz.jl_b200/csrc/sb_los.cu)                                       -> sbl_*
 *   libsbc.so          (model independent, links NCCL; symboltz.jl_b200/csrc/sb_comm.cu)                          -> sbc_*
 *
 * Conventions: plain pointers and sizes only; doubles are IEEE binary64; τ in 1/H0, k in H0/c (reference
 * docs/src/conventions.md:14-18).  Pointers named d* are DEVICE pointers (cudaMalloc / torch / CUDA.jl memory), all others
 * are host pointers (the *_host entry points take host pointers only).  `stream` is a cudaStream_t passed as void* (NULL = default stream); device entry points are
 * asynchronous on that stream.  Return value: 0 (or a non-negative count where stated) on success, negative on error
 * (-1000 - cudaError for CUDA failures).  There is no CPU fallback: without a GPU every device entry point fails.
 *
 * Each entry point names the reference interface it replaces (file:line relative to hersle/SymBoltz.jl v1.6.0).
 */
#ifndef SYMBOLTZ_B200_H
#define SYMBOLTZ_B200_H
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ model library (libsbm_<model>.so) */

/* Model constants.  out[16] = N, NPAR, NBETA, NB, LMAX, NX, W0WA, NNZ(W), FLOPS_F, FLOPS_LU, FLOPS_SOLVE, NLEVELS,
 * NBLOCKS, index of kappa0 in P, index of tau0 in P, NSLOT.   (replaces: `length(unknowns(prob.pt.f.sys))`,
 * sparsity summary of Base.show(::CosmologyProblem), src/solve.jl:40-60) */
int sbm_info(int* out);
const char* sbm_key(void);

/* Parameter vector P[NPAR] (all entry points): h, Omega_c, Omega_b, Omega_g, Omega_nu, C_h = 3/(8π) Omega_h/Iρ0, Omega_L,
 * T0, YHe, fHe, y0, w0, wa, cs2X, kappa0, tau0, then x_i[NX], W_i[NX], dlnf0_i[NX] (momentum quadrature,
 * src/models/neutrinos.jl:55-60,73).  kappa0 and tau0 are filled from sbm_solvebg's info (callback of src/solve.jl:183-189). */

/* Background + thermodynamics solve on the host (replaces solvebg, src/solve.jl:427-435, with the "today" callback of
 * src/solve.jl:158-202 and the spline construction of src/utils.jl:118-127).  Writes the nb Hermite-spline knots
 * t[nb], y[nb][5] = (a, _κ, XH+, XHe+, ΔT), dy[nb][5]; info[0..7] = tau0, kappa0, taurec, retcode, naccept, nreject, length of the
 * solver step in which a crossed 1, 0.
 * Returns nb, or -1 if cap is too small. */
int sbm_solvebg(const double* P, double tini, double tmax, double reltol, double abstol, int cap, double* t, double* y, double* dy, double* info);

/* The same solve in LOCKSTEP with a finished one (parameter lanes of the sensitivity path, BASELINE config 5): takes the knots
 * tfix[nfix] of the primal solve and the length dtlast of its event step (the primal's info[6]) instead of choosing steps, without
 * error control, so that the result is a smooth function of the parameters -- the frozen-step discrete map is what the reference's
 * ForwardDiff duals differentiate (parameters as Duals through solvebg, src/solve.jl:278-284, test/runtests.jl:363-422).
 * info as sbm_solvebg, info[6] = length of the event step, info[7] = 1 if the knot count differs from nfix (lockstep lost). */
int sbm_solvebg_lock(const double* P, double tini, double tmax, int nfix, const double* tfix, double dtlast, int cap, double* t, double* y, double* dy, double* info);

/* The same background solve for n cosmologies in one kernel launch, one thread per cosmology (parameter sweeps: the reference
 * calls solvebg once per θ on the host, docs/src/forecasting.md:56-59 -> src/solve.jl:427-435).
 * dP [n][NPAR] device (the kappa0 and tau0 slots are filled in), dt [n][cap], dy / ddy [n][cap][5], dinfo [n][8] (layout of
 * sbm_solvebg's info, two spare), dnb [n] = knots of each cosmology or -1 if cap was too small.  Asynchronous on `stream`. */
int sbm_solvebg_batch(int n, double* dP, double tini, double tmax, double reltol, double abstol, int cap, double* dt, double* dy, double* ddy, double* dinfo, int* dnb, void* stream);

/* Build the β-table (Jacobian basis functions on the knot-aligned grid) on the device from the uploaded knots.
 * dtab must hold ((nb-1)*msub + 1) * 2 * NBETA doubles.  (replaces the in-RHS spline evaluation `splvalue`,
 * src/utils.jl:135-140, 203-209) */
int sbm_build_table(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, double* dtab, void* stream);

/* Perturbation solve over nk independent k-modes (replaces solvept(ptprob, bgsol, ks, ptivini; reltol, abstol, saveat, ...),
 * src/solve.jl:543-569, i.e. Rodas5P + KLU per mode, src/solve.jl:327-341).
 *   dlut[nlut]: interval look-up (knot interval containing exp(s0 + q*dsl));  dks[nk], dtini[nk] (already clamped to the
 *   background span, src/solve.jl:527); dorder[nk]: processing order (NULL = natural; pass descending k);
 *   dsaveat[nsave] ascending (dense output); dusave[nk][nsave][N] or NULL; duend[nk][N]; dretcode[nk]
 *   (0 Success, 1 MaxIters, 2 DtLessThanMin, 3 Unstable -- SciML retcodes, src/solve.jl:407-419);
 *   dstats[nk][4] = naccept, nreject, nf, nsolve; dqueue: one int of scratch; nctas <= 0: fill the GPU;
 *   dtrace/ntrace: optional (t, dt, EEst) trace of mode 0 (NULL/0 to disable).
 * Returns the launched grid size (>= 0) or a negative error. */
int sbm_solvept(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                const double* dtab, int nk, const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat,
                double reltol, double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, int nctas,
                void* stream, double* dtrace, int ntrace);

/* The same solve under a STATIC, preemptive schedule built by the caller from a per-mode cost estimate (no reference
 * counterpart: the reference spawns one dynamic task per mode, src/solve.jl:566; this replaces the atomic queue when the batch
 * has fewer than ~2 modes per resident warp and the queue's non-preemptive makespan is far from sum/warps).
 * ditems[nitems][3] = (mode, quota, cont): quota > 0 parks the mode after that many attempted steps (at the next accepted
 * step) and publishes a continuation record; cont = 1 waits for the record (bounded: 5 s, then retcode 4) and integrates to
 * the end.  dibeg[nlists + 1]: item range of each warp; nlists must be a multiple of the warps per CTA and at most
 * sbm_resident_warps() (every list has to be resident from the start).  dcont: nk * sbm_cont_stride() doubles, dflags: nk ints.
 * Results are bit-identical to sbm_solvept for every schedule. */
int sbm_solvept_sched(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                      const double* dtab, int nk, const double* dks, const double* dtini, double tend, int nsave, const double* dsaveat, double reltol,
                      double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const int* ditems,
                      const int* dibeg, int nlists, double* dcont, int* dflags, void* stream);
/* One background cosmology as the integrator sees it (device pointers; 128 bytes, natural C layout). */
typedef struct {
    const double* P;              /* [npar] parameters incl. tau0, kappa0 */
    int nb;                       /* spline knots */
    const double *t, *y, *dy;     /* [nb], [nb][5], [nb][5] */
    int tb_nb, msub, nlut;        /* beta-table: knots (= nb), sub-intervals per knot interval, look-up size */
    double s0, inv_dsl;           /* lut[q] = knot interval containing exp(s0 + q / inv_dsl) */
    const double* tb_t;           /* = t */
    const int* lut;               /* [nlut] */
    const double* tab;            /* [(nb-1)*msub + 1][2][NBETA], from sbm_build_table */
    double tend;                  /* end of the integration (tau0) */
    const double* saveat;         /* [nsave] or NULL */
    const double* srcbg;          /* [nsave][sbm_srcbg_stride()] from sbm_srcbg at `saveat` (fused sources, sbm_solvept_batch_src), or NULL */
    double taurec;                /* time of maximal visibility (only the lensing source Spsi uses it) */
} sbm_cosmo_t;

/* Request for the fused evaluation of the CMB source functions inside the perturbation solve (host struct, passed by pointer). */
typedef struct {
    const double* dsrcbg;         /* device [nsave][sbm_srcbg_stride()] from sbm_srcbg (single-cosmology calls; batched: sbm_cosmo_t.srcbg) */
    double* dS;                   /* device output [nk][nS][nsave] */
    int nS, scale_k;              /* nS = 2 (ST, SE) or 3 (+ Spsi); scale_k != 0 returns (k ST, k^2 SE), src/observables/angular.jl:293 */
    double taurec;                /* single-cosmology calls */
} sbm_src_t;

/* Perturbation solve of a BATCH of cosmologies in one launch: mode i belongs to cosmology dcosmo_of[i] (replaces the serial outer
 * loop `for theta: spectrum_matter(probgen(theta), ks)` of docs/src/forecasting.md:56-59 / SURVEY 8b "batched variants with leading
 * ncosmo dimension").  dcosmos: device array of ncosmo sbm_cosmo_t.  dorder (queue order, may be NULL) or a static schedule
 * (ditems/dibeg/nlists/dcont/dflags as in sbm_solvept_sched; all NULL/0 for the queue).  Results per mode are bit-identical to
 * sbm_solvept on that mode's cosmology. */
int sbm_solvept_batch(int ncosmo, const void* dcosmos, int nk, const double* dks, const double* dtini, const int* dcosmo_of, const int* dorder, int nsave, double reltol,
                      double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const int* ditems, const int* dibeg,
                      int nlists, double* dcont, int* dflags, void* stream);
/* The same three solves with the source functions S(tau, k) formed INSIDE the integrator at the save times, from the dense output,
 * and written with coalesced stores (replaces the output_func of solvept in source_grid, src/observables/fourier.jl:267-281, which
 * evaluates getsym(prob.pt, Ss) per save time and keeps only S).  dusave may be NULL: the saved states then never reach HBM.
 * src == NULL: identical to the plain entry points. */
int sbm_solvept_src(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                    const double* dtab, int nk, const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat,
                    double reltol, double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, int nctas,
                    void* stream, const sbm_src_t* src);
int sbm_solvept_sched_src(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                          const double* dtab, int nk, const double* dks, const double* dtini, double tend, int nsave, const double* dsaveat, double reltol,
                          double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const int* ditems,
                          const int* dibeg, int nlists, double* dcont, int* dflags, void* stream, const sbm_src_t* src);
int sbm_solvept_batch_src(int ncosmo, const void* dcosmos, int nk, const double* dks, const double* dtini, const int* dcosmo_of, const int* dorder, int nsave, double reltol,
                          double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const int* ditems, const int* dibeg,
                          int nlists, double* dcont, int* dflags, void* stream, const sbm_src_t* src);
/* Parameter lanes in lockstep ("dual-number lanes in the batched solve", BASELINE config 5; replaces ForwardDiff.Dual parameters pushed
 * through solvept, test/runtests.jl:363-422): G = ncosmo <= 8 neighbouring cosmologies -- lane 0 the primal, lane j >= 1 with one parameter
 * moved by delta_j -- are integrated for the same nk wavenumbers by CTAs of G warps sharing ONE step controller, whose error norm covers
 * the primal and the partials (u^j - u^0) * invdelta[j] (OrdinaryDiffEq's norm of Dual numbers).  With shared, frozen steps
 * (u^j - u^0) / delta_j is the derivative of the discrete solution map, as forward-mode AD gives it, up to O(delta), at the cost of G
 * solves running side by side.  Layout [mode][lane]: dks, dtini, dcosmo_of have nk*G entries (dks[m*G+j] = k_m, dcosmo_of[m*G+j] = j),
 * outputs likewise (duend[nk*G][N], dretcode[nk*G], dstats[nk*G][4], src->dS[nk*G][nS][nsave], dusave optional).  dorder: optional
 * order of the nk groups; invdelta: HOST array of G doubles ([0] unused; 0 leaves a lane's partial out of the norm); tend_common: the
 * smallest sbm_cosmo_t.tend of the lanes.  The lockstep phase runs to tend_common with save decisions on the primal's save times (each
 * lane interpolates at its own); then every lane closes with one private step to its own end time ("today" of its cosmology).
 * Returns the grid size or a negative error. */
int sbm_solvept_lanes(int ncosmo, const void* dcosmos, int nk, const double* dks, const double* dtini, const int* dcosmo_of, const int* dorder, int nsave, double reltol,
                      double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, const double* invdelta, double tend_common,
                      void* stream, const sbm_src_t* src);
/* The single-cosmology solve with ONE CTA per mode ("one CTA per k-mode": SB_R row warps + one warp that evaluates the step controller) instead of one warp per mode: for launches with no
 * more modes than sbm_split_capacity() (BASELINE config 1: 100 modes; the default 61-node C_l path), where the warp-per-mode mapping leaves
 * most of the GPU idle and the run time is the slowest mode's sequential attempts.  The row-parallel phases of an attempt (basis sweep,
 * Jacobian scatter, f-evaluations, stage combinations, error norm, dense output) are spread over the warps, the first solve's three
 * columns go to three warps, path recurrences and the top block stay on one warp while the others fetch and sweep the basis of the later stage
 * times.  8 us instead of 13 us per Rosenbrock attempt on a B200 (profiles/integrate_r2.md).  Arguments and results as sbm_solvept_src (src may be
 * NULL), bit-identical to it.  Returns the grid size, -5 if the model has no split kernel (SB_R outside 2..4), or a negative error.
 * (replaces the same solvept call sites, src/solve.jl:543-569) */
int sbm_split_capacity(void);
int sbm_solvept_split(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                      const double* dtab, int nk, const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat,
                      double reltol, double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream,
                      const sbm_src_t* src);
/* The single-cosmology solve with TRBDF2 instead of Rodas5P: the reference's `ptalg(prob; accuracy = 0)` (src/solve.jl:333-335; solvept's `alg`
 * option, src/solve.jl:543).  Published scheme (Bank et al. 1985, Hosea & Shampine 1996): trapezoidal stage to t + (2 - sqrt 2) dt, BDF2 stage to t + dt,
 * each ONE linear solve with J at the stage time (the system is linear in u), third-order companion error estimate filtered through the last stage's
 * matrix, Gustafsson's predictive step controller, cubic-Hermite dense output; OrdinaryDiffEq.jl's own step selection is not pinned (dependency absent).
 * Arguments and results as sbm_solvept_src (src may be NULL; stats[3] counts linear solves).  One warp per mode, atomic queue.  Returns the grid size
 * or a negative error. */
int sbm_solvept_trbdf2(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                       const double* dtab, int nk, const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat,
                       double reltol, double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream,
                       const sbm_src_t* src);
/* The same with KenCarp4 (`ptalg(prob; accuracy = 1)`, src/solve.jl:336-337; the algorithm of the reference's sparse-Jacobian test, test/runtests.jl:580-590):
 * the ESDIRK half of Kennedy & Carpenter's ARK4(3)6L[2]SA, gamma = 1/4, six stages, stiffly accurate, third-order companion estimate.  Same caveat. */
int sbm_solvept_kencarp4(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                         const double* dtab, int nk, const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat,
                         double reltol, double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream,
                         const sbm_src_t* src);
/* alg: 1 = TRBDF2, 2 = KenCarp4 */
int sbm_solvept_sdirk(int alg, const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int msub, int nlut, double s0, double dsl, const int* dlut,
                       const double* dtab, int nk, const double* dks, const double* dtini, const int* dorder, double tend, int nsave, const double* dsaveat,
                       double reltol, double abstol, int maxiters, double* dusave, double* duend, int* dretcode, long long* dstats, int* dqueue, void* stream,
                       const sbm_src_t* src);
/* Per-save-time background table of the source evaluation at dtaus[nt]: dsrcbg[nt][sbm_srcbg_stride()] = the first three time
 * derivatives of kappa, exp(-kappa), tau0 - tau, 3 spare, beta_m[NBETA], d beta_m/d tau [NBETA] (derivatives along the background
 * flow, as MTK's symbolic expansion of the observed source expressions does, src/solve.jl:637-657). */
int sbm_srcbg(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int nt, const double* dtaus, double* dsrcbg, void* stream);
int sbm_cosmo_bytes(void);
int sbm_cont_stride(void);
int sbm_resident_warps(void);
int sbm_resident_warps_batch(void); /* resident warps of the batched instantiation (its launch bounds are a separate build knob) */
int sbm_warps_per_cta(void);

/* Total-matter gauge-invariant overdensity Δm(τ, k_i) from states du[nk][N] (replaces the observed-function evaluation
 * inside spectrum_matter(sol, k), src/observables/fourier.jl:39-52, 90-97). */
int sbm_delta_m(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, double tau, int nk, const double* dks, const double* du, double* dout, void* stream);

/* CMB source functions at every saved (k, τ) from states kept in HBM -- the stand-alone form of what sbm_solvept_src fuses into the
 * solve; same arithmetic, same bits (replaces getsym(prob.pt, Ss) in source_grid's output_func,
 * src/observables/fourier.jl:267-281, with ST and SE of src/models/cosmologies.jl:99-104).
 * dsrcbg: scratch of nt * sbm_srcbg_stride() doubles.  dS[nk][nS][nt], nS = 2 (ST, SE) or 3 (+ the lensing source Sψ of
 * src/models/cosmologies.jl:105, which needs taurec); scale_k != 0 returns (k·ST, k²·SE) as fed to the line-of-sight integrator
 * (src/observables/angular.jl:293). */
int sbm_srcbg_stride(void);
int sbm_sources(const double* dP, int nb, const double* dt, const double* dy, const double* ddy, int nt, const double* dtaus, double* dsrcbg, int nk,
                const double* dks, const double* dusave, double* dS, int scale_k, int nS, double taurec, void* stream);
int sbm_smem_bytes(void);

/* One-call HOST-buffer variant of the perturbation solve (for hosts without a device allocator: Julia without CUDA.jl, C):
 * every pointer is a host pointer; the library allocates device memory, uploads the nb background knots, builds the β-table
 * and interval look-up (msub = 16, 4096 entries), orders the work queue by descending k, runs sbm_solvept and -- when the
 * outputs are non-NULL -- sbm_delta_m at tend (delta_m[nk]) and the fused source evaluation at the save times (S[nk][nS][nsave],
 * needs nsave > 0; the states are kept in device memory only when usave is asked for), and downloads.  usave[nk][nsave][N], uend[nk][N], retcode[nk], stats[nk][4], delta_m, S may each be NULL.
 * Blocking.  (replaces solvept(ptprob, bgsol, ks, ptivini; saveat, output_func), src/solve.jl:543-569, as called from
 * solve(prob, ks), src/solve.jl:398, spectrum_matter, src/observables/fourier.jl:90-97, and source_grid, fourier.jl:279) */
int sbm_solvept_host(const double* P, int nb, const double* t, const double* y, const double* dy, int nk, const double* ks, const double* tini, double tend, int nsave,
                     const double* saveat, double reltol, double abstol, int maxiters, double* usave, double* uend, int* retcode, long long* stats, double* delta_m, int nS,
                     double taurec, int scale_k, double* S);

/* Host-side diagnostics of the generated code (unit tests of the generator; not a solve path). */
int sbm_debug_fjt(const double* P, const double* y, const double* yp, double tau, double k, const double* u, double* f, double* J, double* dT);
int sbm_debug_split(const double* P, const double* y, const double* yp, double tau, double k, double* Jloc, double* hubs);
int sbm_debug_initial(const double* P, const double* y, double tau, double k, double* u);
int sbm_debug_bg(const double* P, const double* y, double* g, double* J, double* kder, double* obs);
int sbm_debug_delta_m(const double* P, const double* y, double tau, double k, const double* u, double* out);
int sbm_debug_spline(int nb, const double* t, const double* y, const double* dy, double tau, double* yo, double* ypo);
int sbm_debug_beta(const double* P, const double* y, const double* yp, double tau, double* beta, double* betad);

/* ------------------------------------------------------------------ line-of-sight library (libsbl.so) */

/* j_l(x), j_l'(x) table on x = ix*step, ix < nxp, for sorted integer dls[nl]; output dy_[nxp][nl], ddy_[nxp][nl]
 * (l contiguous, = the reference's y[il, ix]).  (replaces SphericalBesselCache, src/observables/angular.jl:18-26, 59-60) */
int sbl_bessel_table(int nl, const int* dls, int nxp, double step, double* dy_, double* ddy_, void* stream);

/* Θ_l(k) for fine-k indices [k0, k0+nk) of dks[nk_total]: optional barycentric interpolation from nc coarse nodes
 * (dBw[nk_total][nc], NULL = sources already on the fine grid), trapezoid weights dwt[nt], χ = dchi[nt], Hermite j_l table,
 * Θ_T/k and Θ_E √((l+2)!/(l−2)!)/k² rescaling; with nS = 3 the third (lensing ψ) source uses the Limber approximation for
 * l >= l_limber (src/observables/angular.jl:155-178).  dSc[nc or nk_total][nS][nt]; dTheta[nS][nl][nk_total].
 * (replaces source_kinterp, src/observables/fourier.jl:232-247; los_integrate, src/observables/angular.jl:109-152;
 * the rescaling of src/observables/angular.jl:301-306) */
int sbl_los(int nk, int k0, int nk_total, const double* dks, int nc, const double* dBw, const double* dSc, int nS, int nt, const double* dchi, const double* dwt, int nl,
            const int* dls, const double* djy, const double* djdy, double invdx, double dx, int nxp, double* dTheta, int l_limber, void* stream);

/* C_l^{AB} = Σ_{k in [k0,k1)} c_k Θ^A_l(k) Θ^B_l(k), c_k = w_k (2/π) k² P0(k) with w_k the natural-cubic-spline integration
 * weights through (0,0) + ks (host-computed).  dCl[nmodes][nl].
 * (replaces spectrum_cmb(ΘlAs, ΘlBs, P0s, ls, ks), src/observables/angular.jl:198-223) */
int sbl_cl(int nl, int nk, int k0, int k1, const double* dck, const double* dTheta, int nmodes, const int* dmodeA, const int* dmodeB, double* dCl, void* stream);

/* Standalone barycentric k-interpolation dSf[nk][n2t] = Σ_j dBw[k][j] dSc[j][n2t] (replaces source_kinterp,
 * src/observables/fourier.jl:232-247) */
int sbl_kinterp(int nk, int nc, const double* dBw, const double* dSc, int n2t, double* dSf, void* stream);

/* FP64 roofline denominator measured on the device the library runs on: independent DFMA chains on every SM, best of `reps`
 * launches of `iters` iterations, result in TFLOP/s through *tflops (host pointer).  Blocking.  (no reference counterpart:
 * measurement support for bench.py's `roofline`, SURVEY 8d "denominator = measured DFMA peak") */
int sbl_dfma_peak(int iters, int reps, double* tflops, void* stream);

/* One-call HOST-buffer variant of k-interpolation + line of sight + C_l: builds the j_l table on the reference's grid
 * (range(0, xmax, length = trunc(xmax/dx)), tabulated up to xcut), uploads ks[nk], Bw[nk][nc] (NULL: Sc is already on the fine
 * grid), Sc[nc or nk][nS][nt], chi[nt], wt[nt], ls[nl], ck[nk], the mode pairs, runs sbl_los + sbl_cl and downloads
 * Cl[nmodes][nl] and, if Theta != NULL, Theta[nS][nl][nk].  Blocking.  Returns -4 if the table (cut at xcut) does not reach
 * max(ks)·max(chi) (the reference asserts jl.x[end] >= kmax·τmax, src/observables/angular.jl:110-116).
 * (replaces the body of spectrum_cmb(modes, prob, jl, ls), src/observables/angular.jl:293-340, after source_grid) */
int sbl_cmb_host(int nk, const double* ks, int nc, const double* Bw, const double* Sc, int nS, int nt, const double* chi, const double* wt, int nl, const int* ls, double dx,
                 double xmax, double xcut, const double* ck, int nmodes, const int* modeA, const int* modeB, int l_limber, double* Cl, double* Theta);

/* ------------------------------------------------------------------ exchange library (libsbc.so; symboltz.jl_b200/csrc/sb_comm.cu) */

/* Multi-GPU exchange steps of the hot path with the NCCL communicator owned by the library (one process or host thread per GPU).
 * (replaces the fan-out over `Threads.@spawn` tasks of src/solve.jl:566 and the serial sweep loop of docs/src/forecasting.md:56-59 across
 * the GPUs of one box: modes / cosmologies are strided over the ranks, results are combined by sum all-reduces over disjoint supports.)
 * Rank 0 obtains the id (sbc_unique_id_bytes() = 128 bytes) and passes it to the other ranks on the host side; sbc_comm_init is
 * collective and binds the communicator to the calling thread's current CUDA device. */
int sbc_unique_id_bytes(void);
int sbc_unique_id(char* out);
int sbc_comm_init(const char* id_bytes, int rank, int world, void** comm);
int sbc_comm_destroy(void* comm);
/* In-place sum over all ranks of the device array dbuf[n], asynchronous on `stream`: (1) gather of the sources S[nk][nS][nt] solved by
 * different ranks (zero-initialised, disjoint supports), (2) the partial C_l sums [nmodes][nl] of the ranks' fine-k slices
 * (sbl_los / sbl_cl with k0, k1), (3) gather of P[ncosmo][nk] in a sharded parameter sweep. */
int sbc_allreduce_sum(void* comm, double* dbuf, long long n, void* stream);
/* The ownership rules of the sharded paths: rank r owns the modes / cosmologies r, r + world, ... (sbc_owned_count of them; the j-th is
 * sbc_owned_index(j, ...)) and the contiguous slice [sbc_slice_begin, sbc_slice_end) of the n fine wavenumbers. */
int sbc_owned_count(int n, int rank, int world);
int sbc_owned_index(int j, int rank, int world);
int sbc_slice_begin(int n, int rank, int world);
int sbc_slice_end(int n, int rank, int world);

#ifdef __cplusplus
}
#endif
#endif
