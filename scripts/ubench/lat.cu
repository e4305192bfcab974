// Micro-benchmarks behind the k_dynamics design: dependent-issue latency of the FP64 pipe, MUFU
// seeds, shared-memory round trips and CTA barriers with 1 warp per scheduler.  nvcc -arch=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_lat(double* out, long long* cyc, int iters) {
    __shared__ double sm[64];
    __shared__ volatile int flag[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double a = 1.0 + 1e-9 * tid, m = 1.0000001, c = 1e-9;
    long long t0, t1;
    // 0: dependent DFMA chain (1 chain)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = fma(a, m, c);
    t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    // 1: two independent DFMA chains
    double b = a + 1;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) { a = fma(a, m, c); b = fma(b, m, c); }
    t1 = clock64();
    if (tid == 0) cyc[1] = t1 - t0;
    // 2: four independent chains
    double d = a + 2, e = a + 3;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) { a = fma(a, m, c); b = fma(b, m, c); d = fma(d, m, c); e = fma(e, m, c); }
    t1 = clock64();
    if (tid == 0) cyc[2] = t1 - t0;
    a += b + d + e;
    // 3: dependent rcp seed + Newton (the vdiv sequence)
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
        r = fma(r, fma(-a, r, 1.0), r);
        const double q = m * r;
        a = fma(fma(-a, q, m), r, q) + 1.0;
    }
    t1 = clock64();
    if (tid == 0) cyc[3] = t1 - t0;
    // 4: shared-memory store -> load round trip (dependent)
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        sm[lane] = a;
        a = ((volatile double*)sm)[lane] + c;
    }
    t1 = clock64();
    if (tid == 0) cyc[4] = t1 - t0;
    // 5: __syncthreads with all warps of the CTA
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        __syncthreads();
        a = fma(a, m, c);
    }
    t1 = clock64();
    if (tid == 0) cyc[5] = t1 - t0;
    // 6: producer/consumer ping-pong between warp 0 and warp 1 through shared memory + named barriers
    if (warp < 2 && blockDim.x >= 64) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (warp == 0) sm[lane] = a;
            asm volatile("bar.sync 1, 64;");
            if (warp == 1) a = sm[lane] + c;
            if (warp == 1) sm[32 + lane] = a;
            asm volatile("bar.sync 2, 64;");
            if (warp == 0) a = sm[32 + lane] + c;
        }
        t1 = clock64();
        if (tid == 0) cyc[6] = t1 - t0;
    }
    // 7: warp shuffle dependent chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + c;
    t1 = clock64();
    if (tid == 0) cyc[7] = t1 - t0;
    // 8: DMUL -> DADD dependent pair
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = (a * m) + c * a;
    t1 = clock64();
    if (tid == 0) cyc[8] = t1 - t0;
    // 9: double select chain (FSEL pair) + DSETP
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = (a > 1.5) ? a * 0.5 : a + 0.25;
    t1 = clock64();
    if (tid == 0) cyc[9] = t1 - t0;
    out[blockIdx.x * blockDim.x + tid] = a;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * 128 * 148);
    cudaMallocManaged(&cyc, sizeof(long long) * 16);
    const int iters = 4096;
    const char* names[] = {"DFMA x1 chain", "DFMA x2 chains (per pair)", "DFMA x4 chains (per quad)", "vdiv sequence + DADD",
                           "STS->LDS round trip", "__syncthreads + DFMA", "2-warp ping-pong (2 named barriers + 2 smem hops)",
                           "SHFL.64 + DADD", "DMUL,DMUL->DFMA", "DSETP+select step"};
    for (int threads : {32, 128}) {
        k_lat<<<1, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        k_lat<<<1, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        printf("threads/CTA %d\n", threads);
        for (int i = 0; i < 10; ++i) printf("  %-52s %8.2f cycles/iter\n", names[i], (double)cyc[i] / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
