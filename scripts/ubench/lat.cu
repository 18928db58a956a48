// Dependent-issue latency micro-benchmarks on one warp (development aid): cycles per dependent operation.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double *out, long long *cyc, double seed) {
    __shared__ double sm[64];
    sm[threadIdx.x] = seed + threadIdx.x;
    sm[threadIdx.x + 32] = 0.5;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-3, y = 1.0000001, z = 1e-9;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) x = fma(x, y, z);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 4 independent DFMA chains
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) { a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    x = a0 + a1 + a2 + a3;
    // 8 independent chains
    double b[8];
#pragma unroll
    for (int j = 0; j < 8; j++) b[j] = x + j;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) b[j] = fma(b[j], y, z);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
#pragma unroll
    for (int j = 0; j < 8; j++) x += b[j];
    // DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) x = x + z;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // dependent LDS chain (pointer chasing through shared memory)
    int idx = threadIdx.x & 31;
    volatile double *vs = sm;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) idx = ((int)vs[idx]) & 31;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // rsqrt chain
    double r = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) r = rsqrt(r) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // shuffle chain (double)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) r += __shfl_xor_sync(0xffffffffu, r, 1 + (i & 15));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // FFMA chain for comparison
    float f = (float)r, g = 1.0001f, h = 1e-6f;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) f = fmaf(f, g, h);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = t1 - t0;
    out[threadIdx.x] = x + r + idx + f;
}
int main() {
    double *out; long long *cyc, h[9];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 9 * 8);
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, cyc, 1.0);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("DFMA dependent: %.1f cyc/op\n4 indep DFMA chains: %.1f cyc/op\n8 indep DFMA chains: %.1f cyc/op\nDADD dependent: %.1f\n"
           "LDS pointer chase (LDS+F2I+LOP): %.1f\nrsqrt+DADD: %.1f\nSHFL(double)+DADD: %.1f\nFFMA dependent: %.1f\nDMUL dependent: %.1f\n",
           h[0] / 512.0, h[1] / 2048.0, h[2] / 4096.0, h[3] / 512.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0, h[7] / 512.0, h[8] / 512.0);
    return 0;
}
