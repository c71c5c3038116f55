// Developer probe: FP64 DMMA (mma.sync m8n8k4) throughput / latency on sm_100a versus warps per SM and independent
// accumulator chains, burst versus sustained, with the SM clock measured inside the kernel (clock64 / globaltimer).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int CH>
__global__ void k_dmma(double *sink, int iters, unsigned long long *clk) {
    double c[CH][2];
#pragma unroll
    for (int q = 0; q < CH; q++) c[q][0] = c[q][1] = 0.0;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    unsigned long long t0 = 0, c0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { t0 = gtime(); c0 = clock64(); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < CH; q++) dmma(c[q][0], c[q][1], a, b);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = clock64() - c0; clk[1] = gtime() - t0; }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < CH; q++) s += c[q][0] + c[q][1];
    if (s == 123.456) sink[0] = s;
}
__global__ void k_dfma(double *sink, int iters, unsigned long long *clk) {
    double c[8];
#pragma unroll
    for (int q = 0; q < 8; q++) c[q] = 1e-3 * q;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
    unsigned long long t0 = 0, c0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { t0 = gtime(); c0 = clock64(); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 8; q++) c[q] = fma(c[q], a, b);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = clock64() - c0; clk[1] = gtime() - t0; }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += c[q];
    if (s == 123.456) sink[0] = s;
}
template <int CH>
void run(int ctas_per_sm, int threads, int iters, int reps, const char *tag) {
    double *sink; unsigned long long *clk, h[2];
    cudaMalloc(&sink, 8); cudaMalloc(&clk, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    k_dmma<CH><<<grid, threads>>>(sink, 16, clk);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) k_dmma<CH><<<grid, threads>>>(sink, iters, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
    const double flop = (double)grid * (threads / 32) * iters * CH * 512.0 * reps;
    const double cyc_per_dmma_smsp = (double)h[0] / ((double)iters * CH * (ctas_per_sm * threads / 32) / 4.0);
    printf("%-10s CH=%d warps/SM=%3d iters=%7d reps=%3d  %8.2f ms  %6.2f TFLOP/s  SM clock %.0f MHz  cycles/DMMA/SMSP %.2f  cycles/iter/warp %.1f\n", tag, CH,
           ctas_per_sm * threads / 32, iters, reps, ms, flop / ms / 1e9, 1e3 * h[0] / (double)h[1], cyc_per_dmma_smsp, (double)h[0] / iters);
    cudaFree(sink); cudaFree(clk);
}
int main() {
    run<8>(8, 256, 4096, 1, "burst");
    run<8>(8, 256, 4096, 4, "burst x4");
    run<8>(8, 256, 4096, 100, "0.4 s");
    run<8>(8, 256, 4096, 500, "2 s");
    run<8>(8, 256, 4096, 4, "after");
    run<1>(1, 128, 100000, 1, "latency");      // one warp per SMSP, one chain: cycles/iter = DMMA latency
    run<2>(1, 128, 100000, 1, "2 chains");
    run<4>(1, 128, 100000, 1, "4 chains");
    run<8>(1, 128, 50000, 1, "8 chains");
    run<9>(1, 128, 50000, 1, "9 chains");
    run<9>(1, 288, 50000, 1, "9w x 9ch");
    run<9>(2, 288, 50000, 1, "18w x 9ch");
    run<6>(1, 256, 50000, 1, "8w x 6ch");
    run<8>(2, 256, 50000, 1, "16w x 8ch");
    {   // DFMA
        double *sink; unsigned long long *clk, h[2];
        cudaMalloc(&sink, 8); cudaMalloc(&clk, 16);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int grid = 148 * 8, iters = 20000;
        k_dfma<<<grid, 256>>>(sink, 16, clk);
        cudaEventRecord(e0);
        for (int r = 0; r < 20; r++) k_dfma<<<grid, 256>>>(sink, iters, clk);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
        printf("DFMA: %.2f TFLOP/s  SM clock %.0f MHz\n", (double)grid * 256 * iters * 8 * 2.0 * 20 / ms / 1e9, 1e3 * h[0] / (double)h[1]);
    }
    return 0;
}
