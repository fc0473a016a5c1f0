// FP64 FMA peak microbenchmark (roofline denominator for the FP64 kernels): independent DFMA chains on all SMs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_peak scripts/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int blocks = nsm * 8, threads = 256, iters = 1 << 16;
    double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        dfma<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        printf("rep %d: %.3f ms  %.2f TFLOP/s FP64\n", rep, ms, tf);
    }
    printf("{\"fp64_tflops\": %.3f, \"sms\": %d}\n", best, nsm);
    return 0;
}
