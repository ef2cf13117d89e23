// Micro-benchmark: latency / throughput of DFMA, I2F.F64 and SHFL on the device (used to size the exact-recompute path).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/micro/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n, int warps_active) {
    double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
    float fa = threadIdx.x * 1e-3f, fb = 1.0000001f, fc = 1e-9f;
    int iv = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) a = fma(a, b, c);
    long long t1 = clock64();
    for (int i = 0; i < n; i++) fa = fmaf(fa, fb, fc);
    long long t2 = clock64();
    double acc = 0;
    for (int i = 0; i < n; i++) { acc += (double)iv; iv = iv * 3 + 1; }
    long long t3 = clock64();
    double s = a;
    for (int i = 0; i < n; i++) s = __shfl_sync(0xffffffffu, s, (threadIdx.x + 1) & 31);
    long long t4 = clock64();
    // 4 independent DFMA chains (throughput)
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3;
    for (int i = 0; i < n; i++) { x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c); }
    long long t5 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + fa + acc + s + x0 + x1 + x2 + x3;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 64);
    const int n = 4096;
    for (int threads : {32, 128, 512}) {
        k<<<148, threads>>>(out, cyc, n, 0); cudaDeviceSynchronize();
        printf("threads/SM %4d: DFMA dep %.1f cyc  FFMA dep %.1f  I2F.F64+DADD dep %.1f  SHFL.f64 dep %.1f  4xDFMA %.1f cyc per 4\n", threads,
               (double)cyc[0] / n, (double)cyc[1] / n, (double)cyc[2] / n, (double)cyc[3] / n, (double)cyc[4] / n);
    }
    return 0;
}
