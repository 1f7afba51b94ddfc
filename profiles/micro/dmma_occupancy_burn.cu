// DMMA (mma.sync.m8n8k4.f64) throughput vs warps per SM and independent accumulator chains per warp
#include <cstdio>
#include <cuda_runtime.h>
template<int CH>
__global__ void burn(double* out, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - threadIdx.x * 1e-4;
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int CH> void run(int blocks_per_sm, int thr, int sms) {
    double* d; cudaMalloc(&d, sizeof(double) * 148 * 64 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096; int blocks = sms * blocks_per_sm; float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); burn<CH><<<blocks, thr>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best; }
    double fl = 2.0 * 256 * CH * iters * (double)blocks * (thr / 32);
    printf("DMMA chains %2d warps/SM %2d : %.2f TFLOP/s\n", CH, blocks_per_sm * thr / 32, fl / (best * 1e-3) * 1e-12);
    cudaFree(d);
}
int main() { int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<1>(1,128,sms); run<2>(1,128,sms); run<4>(1,128,sms); run<8>(1,128,sms);
  run<1>(1,512,sms); run<2>(1,512,sms); run<4>(1,512,sms); run<8>(1,512,sms);
  run<1>(2,1024,sms); run<4>(2,1024,sms);
  return 0; }
