// Single-warp latency microbenchmarks on B200 (dependent DFMA chain, shared-memory load-to-use, warp shuffle, syncwarp).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/lat_bench scripts/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
    __shared__ double sm[64];
    __shared__ int idx[64];
    int lane = threadIdx.x;
    sm[lane] = lane * 0.5; sm[lane + 32] = 1.0; idx[lane] = (lane + 1) & 31; idx[lane + 32] = lane;
    __syncwarp();
    double x = lane;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) x = fma(x, a, b);
    long long t1 = clock64();
    double y = lane; 
    for (int i = 0; i < n; i++) y = y * a;
    long long t2 = clock64();
    int j = lane;
    for (int i = 0; i < n; i++) j = idx[j];          // dependent shared loads (int)
    long long t3 = clock64();
    double z = lane;
    for (int i = 0; i < n; i++) z = __shfl_xor_sync(0xffffffffu, z, 1) + 1.0;   // shuffle + dadd
    long long t4 = clock64();
    double w = 0;
    for (int i = 0; i < n; i++) { sm[lane] = w; __syncwarp(); w = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }  // store -> sync -> load -> sync round trip
    long long t5 = clock64();
    double d = lane + 1.0;
    for (int i = 0; i < n; i++) d = 1.0 / (d + 1.0);   // dependent FP64 division
    long long t6 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
    out[lane] = x + y + j + z + w + d;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 256); cudaMalloc(&c, 64);
    int n = 4096;
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(d, c, 0.9999, 1e-9, n);
    long long h[6]; cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
    const char* names[6] = {"dependent DFMA", "dependent DMUL", "dependent LDS (int)", "SHFL.64 + DADD", "STS->syncwarp->LDS->syncwarp (+DADD)", "dependent FP64 div (+DADD)"};
    for (int i = 0; i < 6; i++) printf("%-40s %.1f cycles per iteration\n", names[i], (double)h[i] / n);
    return 0;
}
