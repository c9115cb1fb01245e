// FP64 peak probe for the roofline denominators of this repo (not product code).
//   - DFMA issue peak, DMMA.8x8x4 issue peak (register-resident operands)
//   - cuBLAS Dgemm (NT) and cuSOLVER Dpotrf as library comparators
//   - device copy bandwidth
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu -lcublas -lcusolver
// Prints one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double seed) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = seed + i + threadIdx.x;
    double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES>
__global__ void dmma_kernel(double* out, int iters, double seed) {
    double c0[TILES], c1[TILES];
#pragma unroll
    for (int i = 0; i < TILES; ++i) { c0[i] = 0; c1[i] = 0; }
    double a = seed + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// layout check: C = A(8x4) * B(4x8) with documented fragment ownership
__global__ void dmma_layout(const double* A, const double* B, double* C) {
    int lane = threadIdx.x;
    double a = A[(lane >> 2) * 4 + (lane & 3)];        // A[row][k], row = lane/4, k = lane%4
    double b = B[(lane & 3) * 8 + (lane >> 2)];        // B[k][n],  k = lane%4, n = lane/4
    double c0 = 0, c1 = 0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    C[(lane >> 2) * 8 + 2 * (lane & 3)] = c0;
    C[(lane >> 2) * 8 + 2 * (lane & 3) + 1] = c1;
}

__global__ void copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);

    // ---- layout check
    {
        double hA[32], hB[32], hC[64], *dA, *dB, *dC;
        for (int i = 0; i < 32; ++i) { hA[i] = 1 + i * 0.5; hB[i] = 2 - i * 0.25; }
        cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dC, 512);
        cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
        dmma_layout<<<1, 32>>>(dA, dB, dC); CK(cudaDeviceSynchronize());
        cudaMemcpy(hC, dC, 512, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) {
            double r = 0; for (int k = 0; k < 4; ++k) r += hA[i * 4 + k] * hB[k * 8 + j];
            maxerr = std::max(maxerr, fabs(r - hC[i * 8 + j]));
        }
        printf(", \"dmma_layout_maxerr\": %.3g", maxerr);
    }

    // ---- DFMA peak
    {
        const int iters = 4096; double best = 0;
        int thr[] = {128, 256, 512, 1024};
        for (int t : thr) {
            dfma_kernel<8><<<sms * 2, t>>>(out, 16, 1.0);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); dfma_kernel<8><<<sms * 2, t>>>(out, iters, 1.0); cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            double fl = 2.0 * 8 * iters * (double)t * sms * 2;
            double tf = fl / (time_ms(e0, e1) * 1e-3) / 1e12;
            best = std::max(best, tf);
        }
        printf(", \"dfma_tflops\": %.2f", best);
    }
    // ---- DMMA peak, sweep warps per SM
    {
        const int iters = 2048;
        int wps[] = {4, 8, 16, 32};
        printf(", \"dmma_tflops\": {");
        bool first = true;
        for (int w : wps) {
            int t = w * 32; if (t > 1024) continue;
            auto run = [&](int tiles) {
                if (tiles == 4) dmma_kernel<4><<<sms, t>>>(out, iters, 1.0);
                else if (tiles == 8) dmma_kernel<8><<<sms, t>>>(out, iters, 1.0);
                else if (tiles == 16) dmma_kernel<16><<<sms, t>>>(out, iters, 1.0);
                else dmma_kernel<32><<<sms, t>>>(out, iters, 1.0);
            };
            int tl[] = {4, 8, 16, 32};
            for (int tiles : tl) {
                run(tiles); CK(cudaDeviceSynchronize());
                cudaEventRecord(e0); run(tiles); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                double fl = 2.0 * 256 * tiles * iters * (double)w * sms;
                double tf = fl / (time_ms(e0, e1) * 1e-3) / 1e12;
                printf("%s\"w%d_t%d\": %.2f", first ? "" : ", ", w, tiles, tf); first = false;
            }
        }
        printf("}");
    }
    // ---- copy bandwidth
    {
        size_t n = (size_t)1 << 28;  // 2^28 doubles = 2 GiB
        double *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
        cudaMemset(a, 1, n * 8); cudaMemset(b, 0, n * 8);
        float best = 1e9;
        for (int r = 0; r < 6; ++r) {
            cudaEventRecord(e0); copy_kernel<<<sms * 16, 512>>>((double2*)a, (double2*)b, n / 2); cudaEventRecord(e1);
            CK(cudaDeviceSynchronize()); if (r) best = std::min(best, time_ms(e0, e1));
        }
        printf(", \"copy_gbs\": %.1f", 2.0 * n * 8 / (best * 1e-3) / 1e9);
        cudaFree(a); cudaFree(b);
    }
    // ---- cuBLAS Dgemm + cuSOLVER Dpotrf
    {
        cublasHandle_t h; cublasCreate(&h);
        cusolverDnHandle_t s; cusolverDnCreate(&s);
        int sizes[] = {2048, 4096, 5632, 8192};
        printf(", \"cublas_dgemm_nt_tflops\": {");
        bool first = true;
        for (int n : sizes) {
            double *A, *B, *C; size_t bytes = (size_t)n * n * 8;
            CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
            std::vector<double> hA((size_t)n * n);
            for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
            cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice); cudaMemcpy(B, hA.data(), bytes, cudaMemcpyHostToDevice);
            double al = 1.0, be = 0.0; float best = 1e9;
            for (int r = 0; r < 5; ++r) {
                cudaEventRecord(e0);
                cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n);
                cudaEventRecord(e1); CK(cudaDeviceSynchronize()); if (r) best = std::min(best, time_ms(e0, e1));
            }
            printf("%s\"%d\": %.2f", first ? "" : ", ", n, 2.0 * n * n * (double)n / (best * 1e-3) / 1e12); first = false;
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        printf("}, \"cusolver_dpotrf\": {");
        first = true;
        for (int n : sizes) {
            size_t bytes = (size_t)n * n * 8; double *A, *A0;
            CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&A0, bytes));
            std::vector<double> hA((size_t)n * n);
            for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                double d = (double)(i - j) / n; hA[(size_t)j * n + i] = exp(-50.0 * d * d) + (i == j ? 0.01 : 0.0);
            }
            cudaMemcpy(A0, hA.data(), bytes, cudaMemcpyHostToDevice);
            int lwork = 0; cusolverDnDpotrf_bufferSize(s, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork);
            double* work; CK(cudaMalloc(&work, sizeof(double) * lwork)); int* info; CK(cudaMalloc(&info, 4));
            float best = 1e9; int hinfo = -1;
            for (int r = 0; r < 4; ++r) {
                cudaMemcpy(A, A0, bytes, cudaMemcpyDeviceToDevice);
                cudaEventRecord(e0);
                cusolverDnDpotrf(s, CUBLAS_FILL_MODE_LOWER, n, A, n, work, lwork, info);
                cudaEventRecord(e1); CK(cudaDeviceSynchronize()); if (r) best = std::min(best, time_ms(e0, e1));
            }
            cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
            double fl = (double)n * n * n / 3.0 + (double)n * n / 2.0;
            printf("%s\"%d\": {\"ms\": %.3f, \"tflops\": %.2f, \"info\": %d}", first ? "" : ", ", n, best, fl / (best * 1e-3) / 1e12, hinfo); first = false;
            cudaFree(A); cudaFree(A0); cudaFree(work); cudaFree(info);
        }
        printf("}");
    }
    printf("}\n");
    return 0;
}
