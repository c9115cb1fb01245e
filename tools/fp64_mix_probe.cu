// FP64 issue probe (not product code): (a) DFMA throughput as a function of resident warps per SM and of the
// independent chains per thread; (b) DFMA warps running next to DMMA warps on the same SM — do the scalar FP64
// instructions fill the pipe bubbles of the DMMA warps, or do the two simply add?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_mix tools/fp64_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__device__ __forceinline__ double dfma_work(int iters, double seed) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = seed + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    return s;
}
__device__ __forceinline__ double dmma_work(int iters, double seed) {
    double c0[16], c1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c0[i] = 0; c1[i] = 0; }
    double a = seed, b = 1.0 - seed * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
    return s;
}
template <int CHAINS>
__global__ void dfma_only(double* out, int iters) { out[blockIdx.x * blockDim.x + threadIdx.x] = dfma_work<CHAINS>(iters, threadIdx.x * 1e-3); }
// warps [0, ndmma) run DMMA (iters_dmma x 16), the rest run DFMA (iters_dfma x 8 chains)
__global__ void mix(double* out, int ndmma, int iters_dmma, int iters_dfma) {
    const int warp = threadIdx.x >> 5;
    double s = (warp < ndmma) ? dmma_work(iters_dmma, threadIdx.x * 1e-3) : dfma_work<8>(iters_dfma, threadIdx.x * 1e-3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
static float run(void (*launch)(), int reps = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = best < ms ? best : ms;
    }
    return best;
}
static double* g_out; static int g_sms, g_threads, g_it1, g_it2, g_nd;
template <int C> static void l_dfma() { dfma_only<C><<<g_sms, g_threads>>>(g_out, g_it1); }
static void l_mix() { mix<<<g_sms, g_threads>>>(g_out, g_nd, g_it1, g_it2); }
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); g_sms = p.multiProcessorCount;
    cudaMalloc(&g_out, sizeof(double) * g_sms * 1024);
    printf("{\"dfma_tflops\": {");
    bool first = true;
    for (int warps : {4, 8, 12, 16, 32}) {
        g_threads = warps * 32; g_it1 = 200000;
        float t4 = run(l_dfma<4>), t8 = run(l_dfma<8>);
        double f4 = 2.0 * 4 * g_it1 * (double)g_threads * g_sms / (t4 * 1e-3) / 1e12, f8 = 2.0 * 8 * g_it1 * (double)g_threads * g_sms / (t8 * 1e-3) / 1e12;
        printf("%s\"w%d_c4\": %.2f, \"w%d_c8\": %.2f", first ? "" : ", ", warps, f4, warps, f8); first = false;
    }
    printf("}, \"mix_8dmma_plus_dfma_warps\": {");
    // 8 DMMA warps alone, then with 4 / 8 extra DFMA warps whose work is 25 % of the DMMA pipe time
    first = true;
    for (int extra : {0, 4, 8}) {
        g_nd = 8; g_threads = (8 + extra) * 32; g_it1 = 20000;               // 8 warps x 20000 x 16 DMMA = 16 clk each -> 2 per SMSP
        // DMMA pipe time per SMSP: 2 warps x 20000 x 16 x 16 clk = 10.24 M clk; DFMA: w warps/SMSP x iters x 8 x 2 clk
        g_it2 = extra ? (int)(0.25 * 2 * 20000 * 16 * 16 / ((extra / 4.0) * 8 * 2)) : 0;
        float t = run(l_mix);
        double dm = 8.0 * 32 * g_sms * (double)g_it1 * 16 * 512 / 32 / (t * 1e-3) / 1e12;     // DMMA flops: 512 per warp instr
        double df = extra ? 2.0 * 8 * g_it2 * (double)(extra * 32) * g_sms / (t * 1e-3) / 1e12 : 0.0;
        printf("%s\"extra%d\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"sum\": %.2f}", first ? "" : ", ", extra, t, dm, df, dm + df); first = false;
    }
    printf("}}\n");
    return 0;
}
