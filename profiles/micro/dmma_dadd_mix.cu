// Does interleaving DADD with DMMA (mma.sync.m8n8k4.f64) cost more FP64-pipe time than the sum of the two? The 3M block path
// of the executor issues, per batch of 8 items in the backward sweep, 18 DMMA and 28 DADD, finely interleaved by the compiler.
// Variants per loop iteration (per warp):  A: 18 DMMA;  D: 28 DADD;  B: both, one DADD (or two) after every DMMA;
// C: both, grouped (8 DADD, 18 DMMA, 20 DADD).  Reports cycles per iteration per SM sub-partition at 4 warps per sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_dadd_mix dmma_dadd_mix.cu && ./dmma_dadd_mix
#include <cstdio>
#include <cuda_runtime.h>

#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))
#define DADD(d, a, b) asm volatile("add.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))

template <int VAR>
__global__ void __launch_bounds__(512, 1) mix(double* out, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - threadIdx.x * 1e-4;
    double c[9][2];
    double s[14];
#pragma unroll
    for (int i = 0; i < 9; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 14; ++i) s[i] = i * 0.5;
    for (int it = 0; it < iters; ++it) {
        if (VAR == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 9; ++i) DMMA(c[i][0], c[i][1], a, b);
        } else if (VAR == 1) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 14; ++i) DADD(s[i], s[i], a);
        } else if (VAR == 2) {  // fine interleave: DMMA, DADD, DMMA, DADD, DADD, ...
            int d = 0;
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    DMMA(c[i][0], c[i][1], a, b);
                    DADD(s[d % 14], s[d % 14], a); ++d;
                    if ((i & 1) == 0 || i == 8) { DADD(s[d % 14], s[d % 14], a); ++d; }
                }
        } else {  // grouped: 8 DADD, 18 DMMA, 20 DADD
#pragma unroll
            for (int i = 0; i < 8; ++i) DADD(s[i], s[i], a);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 9; ++i) DMMA(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int i = 0; i < 20; ++i) DADD(s[i % 14], s[i % 14], a);
        }
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) t += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 14; ++i) t += s[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int VAR>
void run(const char* name, int thr, int sms, double ghz) {
    double* d;
    cudaMalloc(&d, sizeof(double) * 148 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        mix<VAR><<<sms, thr>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    const double warps_per_sp = thr / 32 / 4.0;
    const double clk_per_iter_per_sp = best * 1e-3 * ghz * 1e9 / iters;  // all warps of a sub-partition together
    printf("%-28s warps/SMSP %.0f : %8.1f clk per iteration round (%.1f clk per warp-iteration)\n", name, warps_per_sp, clk_per_iter_per_sp,
           clk_per_iter_per_sp / warps_per_sp);
    cudaFree(d);
}

int main() {
    int sms, khz;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("SMs %d, clock %.3f GHz (nominal)\n", sms, ghz);
    for (int thr : {128, 256, 512}) {
        run<0>("A: 18 DMMA", thr, sms, ghz);
        run<1>("D: 28 DADD", thr, sms, ghz);
        run<2>("B: 18 DMMA + 28 DADD mixed", thr, sms, ghz);
        run<3>("C: 8 DADD,18 DMMA,20 DADD", thr, sms, ghz);
    }
    return 0;
}
