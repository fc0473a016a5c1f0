// symboltz.jl_b200 -- model-independent line-of-sight / C_l kernels for sm_100a + C ABI (device pointers).
//
// Replaces, for the hot path (file:line relative to the reference tree):
//   SphericalBesselCache           src/observables/angular.jl:9-48, 59-60   -> sbl_bessel_table (table built on the GPU)
//   source_kinterp (Chebyshev)     src/observables/fourier.jl:232-247, 524-547 -> fused into sbl_los (B·S_coarse in shared memory)
//   los_integrate                  src/observables/angular.jl:109-152       -> sbl_los (one CTA per fine k, threads over l × τ-slices)
//   Θ rescale                      src/observables/angular.jl:301-306       -> epilogue of sbl_los
//   spectrum_cmb(ΘA,ΘB,P0,ls,ks)   src/observables/angular.jl:198-223       -> sbl_cl (natural-spline k-integral as a weighted dot product;
//                                                                              the integral is linear in the data, weights from the host)
// FP64 throughout; no tensor cores (this is table interpolation + reductions, not a contraction worth reshaping).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#define SBL_CUDA_CHECK(x)                                            \
    do {                                                             \
        cudaError_t e_ = (x);                                        \
        if (e_ != cudaSuccess) { fprintf(stderr, "symboltz_b200(los): CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -(int)e_ - 1000; } \
    } while (0)

#define SBL_MAXNL 1024

// j_l(x) and j_l'(x) = l/(2l+1) j_{l-1} − (l+1)/(2l+1) j_{l+1} for the requested (sorted, integer) l at x = ix·step.
// Upward recurrence where it is stable (l < x), Miller's downward recurrence with rescaling otherwise.
// Output layout: y[ix*nl + il] (l contiguous, as the reference's `y[il, ix]` column-major matrix).
__global__ void sbl_bessel_kernel(int nl, const int* __restrict__ ls, int nxp, double step, double* __restrict__ y, double* __restrict__ dy) {
    int ix = blockIdx.x * blockDim.x + threadIdx.x;
    if (ix >= nxp) return;
    const double x = ix * step;
    double* yo = y + (size_t)ix * nl;
    double* dyo = dy + (size_t)ix * nl;
    const int lmax = ls[nl - 1];
    if (x == 0.0) {
        for (int i = 0; i < nl; i++) { yo[i] = (ls[i] == 0) ? 1.0 : 0.0; dyo[i] = (ls[i] == 1) ? 1.0 / 3.0 : 0.0; }
        return;
    }
    double sx, cx;
    sincos(x, &sx, &cx);
    const double j0 = sx / x, j1 = sx / (x * x) - cx / x;
    if (x > lmax + 1.5) { // upward, stable for n < x
        double jm = j0, jn = j1; // j_{n-1}, j_n with n = 1
        int ip = 0;
        if (ls[0] == 0) { yo[0] = j0; dyo[0] = -j1; ip = 1; }
        for (int n = 1; n <= lmax; n++) {
            double jp = (2 * n + 1) / x * jn - jm; // j_{n+1}
            if (ip < nl && ls[ip] == n) { yo[ip] = jn; dyo[ip] = (n * jm - (n + 1) * jp) / (2 * n + 1); ip++; }
            jm = jn; jn = jp;
        }
        return;
    }
    // downward (Miller).  Start far enough above max(lmax+1, x) that the seed error has decayed below 1e-17.
    const double big = 1e200, small = 1e-200;
    int Lm = lmax + 1;
    int nstart = Lm + 25 + (int)ceil(pow(60.0 * sqrt(0.5 * Lm + 1.0), 2.0 / 3.0));
    double jp = 0.0, jn = 1e-250; // j_{n+1}, j_n at n = nstart (unnormalised)
    int nres = 0;                 // number of rescalings applied so far
    short cnt[SBL_MAXNL];
    int ip = nl - 1;
    for (int n = nstart; n >= 1; n--) {
        double jm = (2 * n + 1) / x * jn - jp; // j_{n-1}
        if (ip >= 0 && ls[ip] == n) { yo[ip] = jn; dyo[ip] = (n * jm - (n + 1) * jp) / (2 * n + 1); cnt[ip] = (short)nres; ip--; }
        jp = jn; jn = jm;
        if (fabs(jn) > big) { jn *= small; jp *= small; nres++; }
    }
    // now jn = j_0, jp = j_1 (unnormalised)
    if (ip >= 0 && ls[ip] == 0) { yo[ip] = jn; dyo[ip] = -jp; cnt[ip] = (short)nres; ip--; }
    const double scale = (fabs(j0) >= fabs(j1)) ? j0 / jn : j1 / jp;
    for (int i = 0; i < nl; i++) {
        int m = nres - cnt[i];
        double f = scale;
        // values recorded before m later rescalings are too large by big^m relative to the final scale
        if (m >= 2) f = 0.0; else if (m == 1) f *= small;
        yo[i] *= f; dyo[i] *= f;
    }
}

// One CTA per fine k: optional barycentric k-interpolation of the coarse sources (in shared memory), trapezoid-weighted
// line-of-sight sum against the Hermite-interpolated j_l table, Θ rescaling.  Threads: x → l, y → τ slice.
// nS sources (2: T, E; 3: + lensing ψ).  Sc layout [nc][nS][nt];  Bw [nk][nc] barycentric weights (NULL: direct, nc == nk);
// jy/jdy [nx][nl];  Theta out [nS][nl][nk].  For the ψ source and l >= l_limber the Limber approximation
// Θ_l = √(π/(2l+1)) S(τ0 − (l+½)/k)/k replaces the integral (reference src/observables/angular.jl:155-178).
#define SBL_MAXS 3
__global__ void sbl_los_kernel(int nk, int k0, const double* __restrict__ ks, int nc, const double* __restrict__ Bw, const double* __restrict__ Sc, int nS, int nt,
                               const double* __restrict__ chi, const double* __restrict__ wt, int nl, const int* __restrict__ ls, const double* __restrict__ jy,
                               const double* __restrict__ jdy, double invdx, double dxc, int nxp, double* __restrict__ Theta, int nk_total, int l_limber) {
    extern __shared__ double sh[];
    double* Sw = sh;                         // [nS][nt] trapezoid-weighted sources
    double* Sraw = sh + nS * nt;             // [nt] unweighted ψ source (Limber), only if nS == 3
    double* red = sh + (nS + (nS > 2 ? 1 : 0)) * nt; // [TS][LT][nS] reduction scratch
    const int ik = blockIdx.x;
    const int LT = blockDim.x, TS = blockDim.y;
    const int tid = threadIdx.y * LT + threadIdx.x, nthr = LT * TS;
    const double k = ks[k0 + ik];
    for (int o = tid; o < nS * nt; o += nthr) {
        int s = o / nt, it = o % nt;
        double v;
        if (Bw) {
            const double* bw = Bw + (size_t)(k0 + ik) * nc;
            v = 0;
            for (int j = 0; j < nc; j++) { const double sc = (it == nt - 1) ? 0.0 : Sc[((size_t)j * nS + s) * nt + it]; v += bw[j] * sc; } // last row (χ = 0) zeroed, angular.jl:296
        } else v = (it == nt - 1) ? 0.0 : Sc[((size_t)(k0 + ik) * nS + s) * nt + it];
        if (s == 2) Sraw[it] = v;
        sh[o] = v * wt[it];
    }
    __syncthreads();
    const int il = threadIdx.x;
    double acc[SBL_MAXS] = {0, 0, 0};
    if (il < nl) {
        for (int it = threadIdx.y; it < nt; it += TS) {
            double w = k * chi[it] * invdx;
            int i = (int)w; // trunc, x >= 0
            w -= i;
            if (i > nxp - 2) { i = nxp - 2; w = 1.0; }
            const double wm1 = w - 1.0;
            const size_t o0 = (size_t)i * nl + il, o1 = o0 + nl;
            const double ym = __ldg(jy + o0), yp = __ldg(jy + o1), dm = __ldg(jdy + o0), dp = __ldg(jdy + o1);
            const double j = (1 + 2 * w) * wm1 * wm1 * ym + w * w * (3 - 2 * w) * yp + w * wm1 * (wm1 * dm + w * dp) * dxc; // angular.jl:38-48
#pragma unroll
            for (int s = 0; s < SBL_MAXS; s++) if (s < nS) acc[s] += Sw[s * nt + it] * j;
        }
    }
    for (int s = 0; s < nS; s++) red[(threadIdx.y * LT + threadIdx.x) * nS + s] = acc[s];
    __syncthreads();
    if (threadIdx.y == 0 && il < nl) {
        for (int t = 1; t < TS; t++) for (int s = 0; s < nS; s++) acc[s] += red[(t * LT + il) * nS + s];
        const double l = (double)ls[il];
        Theta[((size_t)0 * nl + il) * nk_total + k0 + ik] = acc[0] / k;                                                   // angular.jl:302
        Theta[((size_t)1 * nl + il) * nk_total + k0 + ik] = acc[1] * sqrt((l + 2) * (l + 1) * l * (l - 1)) / (k * k);     // angular.jl:305
        if (nS > 2) {
            double th = acc[2];
            if (ls[il] >= l_limber) { // Limber: cubic Hermite in χ with finite-difference slopes, as the reference
                th = 0.0;
                const double chiL = (l + 0.5) / k;
                if (chiL <= chi[0]) {
                    int lo = 0, hi = nt - 1; // first index with χ_i <= χL (χ descending) == searchsortedfirst(τs, τ0 − χL)
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (chi[mid] <= chiL) hi = mid; else lo = mid + 1; }
                    const int im = lo;
                    if (im > 0) {
                        const int ip = im - 1;
                        const double Sm = Sraw[im], Sp = Sraw[ip], chim = chi[im], chip = chi[ip], dchi = chip - chim;
                        const double dSm = (im <= nt - 2) ? (Sraw[im + 1] - Sp) / (chi[im + 1] - chip) : (Sp - Sm) / dchi;
                        const double dSp = (ip >= 1) ? (Sm - Sraw[im - 2]) / (chim - chi[im - 2]) : (Sp - Sm) / dchi;
                        const double t = (chiL - chim) / dchi, t2 = t * t, t3 = t2 * t;
                        const double Sv = (2 * t3 - 3 * t2 + 1) * Sm + (t3 - 2 * t2 + t) * dchi * dSm + (-2 * t3 + 3 * t2) * Sp + (t3 - t2) * dchi * dSp;
                        th = sqrt(3.14159265358979323846 / (2 * l + 1)) * Sv / k;
                    }
                }
            }
            Theta[((size_t)2 * nl + il) * nk_total + k0 + ik] = th;
        }
    }
}

// C_l^{AB} = Σ_k c_k Θ^A_l(k) Θ^B_l(k) with c_k = w_k (2/π) k² P0(k), w_k = natural-cubic-spline integration weights.
// grid (nl, nmodes); modes: pairs (A,B) of source indices.  Cl out [nmodes][nl].  Deterministic tree reduction.
__global__ void sbl_cl_kernel(int nl, int nk, int k0, int k1, const double* __restrict__ ck, const double* __restrict__ Theta, const int* __restrict__ modeA,
                              const int* __restrict__ modeB, double* __restrict__ Cl) {
    __shared__ double red[256];
    const int il = blockIdx.x, im = blockIdx.y;
    const double* TA = Theta + ((size_t)modeA[im] * nl + il) * nk;
    const double* TB = Theta + ((size_t)modeB[im] * nl + il) * nk;
    double acc = 0;
    for (int ik = k0 + threadIdx.x; ik < k1; ik += blockDim.x) acc += ck[ik] * TA[ik] * TB[ik];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) Cl[(size_t)im * nl + il] = red[0];
}

// Standalone barycentric k-interpolation (reference source_kinterp, fourier.jl:232-247): Sf[nk][2][nt] = Σ_j Bw[k][j] Sc[j][2][nt]
__global__ void sbl_kinterp_kernel(int nk, int nc, const double* __restrict__ Bw, const double* __restrict__ Sc, int n2t, double* __restrict__ Sf) {
    int ik = blockIdx.x;
    for (int o = threadIdx.x; o < n2t; o += blockDim.x) {
        double v = 0;
        for (int j = 0; j < nc; j++) v += Bw[(size_t)ik * nc + j] * Sc[(size_t)j * n2t + o];
        Sf[(size_t)ik * n2t + o] = v;
    }
}

extern "C" {

int sbl_bessel_table(int nl, const int* dls, int nxp, double step, double* dy_, double* ddy_, void* stream) {
    if (nl > SBL_MAXNL) return -2;
    sbl_bessel_kernel<<<(nxp + 63) / 64, 64, 0, (cudaStream_t)stream>>>(nl, dls, nxp, step, dy_, ddy_);
    SBL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Line-of-sight integration for fine-k indices [k0, k0+nk) of a grid with nk_total points (multi-GPU: each rank its slice).
int sbl_los(int nk, int k0, int nk_total, const double* dks, int nc, const double* dBw, const double* dSc, int nS, int nt, const double* dchi, const double* dwt, int nl,
            const int* dls, const double* djy, const double* djdy, double invdx, double dx, int nxp, double* dTheta, int l_limber, void* stream) {
    if (nk <= 0) return 0;
    if (nS < 2 || nS > SBL_MAXS) return -3;
    int LT = ((nl + 31) / 32) * 32;
    if (LT > 1024) return -2;
    int TS = 512 / LT; if (TS > 8) TS = 8; if (TS < 1) TS = 1;
    size_t smem = (size_t)((nS + (nS > 2 ? 1 : 0)) * nt + nS * LT * TS) * sizeof(double);
    if (smem > 48 * 1024) SBL_CUDA_CHECK(cudaFuncSetAttribute(sbl_los_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sbl_los_kernel<<<nk, dim3(LT, TS), smem, (cudaStream_t)stream>>>(nk, k0, dks, nc, dBw, dSc, nS, nt, dchi, dwt, nl, dls, djy, djdy, invdx, dx, nxp, dTheta, nk_total, l_limber);
    SBL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sbl_cl(int nl, int nk, int k0, int k1, const double* dck, const double* dTheta, int nmodes, const int* dmodeA, const int* dmodeB, double* dCl, void* stream) {
    sbl_cl_kernel<<<dim3(nl, nmodes), 256, 0, (cudaStream_t)stream>>>(nl, nk, k0, k1, dck, dTheta, dmodeA, dmodeB, dCl);
    SBL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// FP64 roofline denominator, measured in the run that reports it (bench.py): independent DFMA chains on every SM.
// Returns the best of `reps` timed launches in TFLOP/s through *tflops (events on `stream`); dscratch: blocks*256 doubles.
__global__ void sbl_dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int sbl_dfma_peak(int iters, int reps, double* tflops, void* stream) {
    int dev, nsm;
    SBL_CUDA_CHECK(cudaGetDevice(&dev));
    SBL_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = nsm * 8, threads = 256;
    double* d = nullptr;
    SBL_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    SBL_CUDA_CHECK(cudaEventCreate(&e0)); SBL_CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0;
    for (int r = 0; r <= reps; r++) { // launch 0 is the warm-up
        cudaEventRecord(e0, (cudaStream_t)stream);
        sbl_dfma_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, (cudaStream_t)stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    SBL_CUDA_CHECK(cudaGetLastError());
    *tflops = best;
    return 0;
}

int sbl_kinterp(int nk, int nc, const double* dBw, const double* dSc, int n2t, double* dSf, void* stream) {
    if (nk <= 0) return 0;
    sbl_kinterp_kernel<<<nk, 256, 0, (cudaStream_t)stream>>>(nk, nc, dBw, dSc, n2t, dSf);
    SBL_CUDA_CHECK(cudaGetLastError());
    return 0;
}
} // extern "C"

// ---------------------------------------------------------------------------------------------- host-buffer entry point
// One-call variant for hosts without a CUDA allocator: HOST pointers only.  Builds the j_l table, uploads the sources and weights, runs k-interpolation + line of sight + C_l, downloads C_l (and Θ_l(k) if asked).
#include <algorithm>
#include <vector>
struct SblDevBuf {
    std::vector<void*> ptrs;
    ~SblDevBuf() { for (void* q : ptrs) cudaFree(q); }
    template <class T> cudaError_t get(T** out, size_t n, const T* host = nullptr) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(q);
        *out = (T*)q;
        if (host && n) e = cudaMemcpy(q, host, n * sizeof(T), cudaMemcpyHostToDevice);
        return e;
    }
};

extern "C" int sbl_cmb_host(int nk, const double* ks, int nc, const double* Bw, const double* Sc, int nS, int nt, const double* chi, const double* wt, int nl, const int* ls, double dx,
                            double xmax, double xcut, const double* ck, int nmodes, const int* modeA, const int* modeB, int l_limber, double* Cl, double* Theta) {
    if (nk <= 0 || nl <= 0 || nt <= 0 || nmodes <= 0 || !(dx > 0)) return -1;
    // the reference's grid range(0, xmax, length = trunc(xmax/dx)) + one padded point (src/observables/angular.jl:18-25), cut at xcut
    const int n = (int)(xmax / dx);
    if (n < 2) return -1;
    const double step = xmax / (n - 1);
    int nxp = n + 1;
    if (xcut < xmax) nxp = std::min(nxp, (int)ceil(xcut / step) + 2);
    { // the reference asserts jl.x[end] >= kmax·τmax (src/observables/angular.jl:110-116); the kernel would clamp silently
        double kmax = 0, chimax = 0;
        for (int i = 0; i < nk; i++) kmax = std::max(kmax, ks[i]);
        for (int i = 0; i < nt; i++) chimax = std::max(chimax, chi[i]);
        if ((nxp - 2) * step < kmax * chimax) return -4;
    }
    SblDevBuf B;
    double *dks, *dBw = nullptr, *dSc, *dchi, *dwt, *djy, *djdy, *dck, *dTh, *dCl;
    int *dls, *dmA, *dmB;
    SBL_CUDA_CHECK(B.get(&dks, nk, ks));
    if (Bw) SBL_CUDA_CHECK(B.get(&dBw, (size_t)nk * nc, Bw));
    SBL_CUDA_CHECK(B.get(&dSc, (size_t)(Bw ? nc : nk) * nS * nt, Sc)); SBL_CUDA_CHECK(B.get(&dchi, nt, chi)); SBL_CUDA_CHECK(B.get(&dwt, nt, wt));
    SBL_CUDA_CHECK(B.get(&dls, nl, ls)); SBL_CUDA_CHECK(B.get(&djy, (size_t)nxp * nl)); SBL_CUDA_CHECK(B.get(&djdy, (size_t)nxp * nl));
    SBL_CUDA_CHECK(B.get(&dck, nk, ck)); SBL_CUDA_CHECK(B.get(&dTh, (size_t)nS * nl * nk)); SBL_CUDA_CHECK(B.get(&dCl, (size_t)nmodes * nl));
    SBL_CUDA_CHECK(B.get(&dmA, nmodes, modeA)); SBL_CUDA_CHECK(B.get(&dmB, nmodes, modeB));
    int rc = sbl_bessel_table(nl, dls, nxp, step, djy, djdy, nullptr);
    if (rc < 0) return rc;
    rc = sbl_los(nk, 0, nk, dks, Bw ? nc : 0, dBw, dSc, nS, nt, dchi, dwt, nl, dls, djy, djdy, 1.0 / step, dx, nxp, dTh, l_limber, nullptr);
    if (rc < 0) return rc;
    rc = sbl_cl(nl, nk, 0, nk, dck, dTh, nmodes, dmA, dmB, dCl, nullptr);
    if (rc < 0) return rc;
    SBL_CUDA_CHECK(cudaDeviceSynchronize());
    SBL_CUDA_CHECK(cudaMemcpy(Cl, dCl, (size_t)nmodes * nl * sizeof(double), cudaMemcpyDeviceToHost));
    if (Theta) SBL_CUDA_CHECK(cudaMemcpy(Theta, dTh, (size_t)nS * nl * nk * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
