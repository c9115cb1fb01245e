// gemm_tma.cuh — persistent, batched FP64 tensor-core GEMM for K-CONTIGUOUS operands, fed by TMA:
//     C[m][n] (+)= alpha * sum_{k in [klo, khi)} A[m][k] * B[n][k]          (both operands row-major with k contiguous)
// with an optional TRANSPOSED second store CT[n][m] = alpha * (...) — which is what lets every O(n^3) step of the conditioning
// and of the batched marginal likelihood keep its operands k-contiguous (the triangular inverse carries X and X^T, the product
// of the inverse factors reads X^T twice), so that all of them run on the TMA + mbarrier + DMMA pipeline of the sweep kernel
// instead of the cp.async kernels whose MN-contiguous shared-memory layout costs bank conflicts (ncu, round 1: 84.7 % DMMA).
//
// One CTA per SM, static serpentine schedule over (batch item, tile) work items ordered heaviest first; the producer warp runs
// ahead across item boundaries (no pipeline fill / drain per tile — the short k-loops of the batched n ~ 1000 matrices are
// exactly where that matters); k-ranges that are structurally zero are skipped (flags as in gemm_dmma.cuh).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm_dmma.cuh"
#include "sweep_tma.cuh"

namespace abo {

struct TmaGemmParams {
    int Mt, Nt, batch, total;                       // tiles per item, items, work items in this launch
    int K;                                          // full k extent (multiple of 16)
    int flags;                                      // KLO_M | KLO_N | KHI_M | LOWER_ONLY
    // tile origins in the 2-D tensor maps (k coordinate, row coordinate), per batch item z:
    int a_row0, a_rstep, a_col0, a_cstep;           // A rows  a_row0 + z a_rstep + m0,  k columns a_col0 + z a_cstep + k
    int b_row0, b_rstep, b_col0, b_cstep;
    double* C;  int64_t ldc,  c_off0,  c_zstep;     // C  element (m, n): C [c_off0  + z c_zstep  + m ldc  + n]   (nullable)
    double* CT; int64_t ldct, ct_off0, ct_zstep;    // CT element (n, m): CT[ct_off0 + z ct_zstep + n ldct + m]   (nullable, beta = 0)
    double alpha, beta;
};

struct TgTile { int z, m0, n0, klo, nk; };
__device__ __forceinline__ TgTile tg_decode(const TmaGemmParams& p, int t) {
    TgTile w;
    w.z = t % p.batch;
    const int tl = t / p.batch;
    int mt, nt;
    if (p.flags & LOWER_ONLY) {                     // lower tiles nt <= min(mt, Nt - 1) of an Mt x Nt grid (Nt <= Mt): the triangle of
                                                    // the first Nt tile rows, then full rows of Nt tiles (a trapezoid when Nt < Mt)
        const int tri = p.Nt * (p.Nt + 1) / 2;
        if (tl < tri) {
            mt = (int)((sqrt(8.0 * tl + 1.0) - 1.0) * 0.5);
            while ((mt + 1) * (mt + 2) / 2 <= tl) ++mt;
            while (mt * (mt + 1) / 2 > tl) --mt;
            nt = tl - mt * (mt + 1) / 2;
        } else { mt = p.Nt + (tl - tri) / p.Nt; nt = (tl - tri) % p.Nt; }
    } else if (p.flags & KLO_N) { nt = tl / p.Mt; mt = tl % p.Mt; }
    else if (p.flags & KHI_M) { mt = p.Mt - 1 - tl / p.Nt; nt = tl % p.Nt; }
    else { mt = tl / p.Nt; nt = tl % p.Nt; }
    w.m0 = mt * 128; w.n0 = nt * 128;
    int klo = 0, khi = p.K;
    if (p.flags & KLO_M) klo = w.m0;
    if ((p.flags & KLO_N) && w.n0 > klo) klo = w.n0;
    if ((p.flags & KHI_M) && w.m0 + 128 < khi) khi = w.m0 + 128;
    w.klo = klo;
    w.nk = (khi > klo) ? (khi - klo) / 16 : 0;
    return w;
}

constexpr int TG_SMEM_BYTES = SW_STAGES * SW_STAGE_BYTES + 2 * SW_STAGES * 8 + 128;

__global__ void __launch_bounds__(SW_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TmaGemmParams p) {
    extern __shared__ __align__(128) unsigned char tg_smem[];
    double* stage_base = reinterpret_cast<double*>(tg_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(tg_smem + SW_STAGES * SW_STAGE_BYTES);
    uint64_t* empty = full + SW_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < SW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // work item r of this CTA: serpentine over the launch's item list
    auto item_of = [&](int r) { return r * G + ((r & 1) ? (G - 1 - b) : b); };

    if (warp == 8) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int r = 0;; ++r) {
                const int t = item_of(r);
                if (t >= p.total) { if (r * G >= p.total) break; else continue; }
                const TgTile w = tg_decode(p, t);
                const int arow = p.a_row0 + w.z * p.a_rstep + w.m0, acol = p.a_col0 + w.z * p.a_cstep + w.klo;
                const int brow = p.b_row0 + w.z * p.b_rstep + w.n0, bcol = p.b_col0 + w.z * p.b_cstep + w.klo;
                for (int kt = 0; kt < w.nk; ++kt) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], SW_STAGE_BYTES);
                    double* sa = stage_base + (size_t)stage * (2 * SW_OPER_DOUBLES);
                    double* sb = sa + SW_OPER_DOUBLES;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        tma_load_2d(sa + q * 512, &tmA, acol + kt * 16 + 4 * q, arow, &full[stage]);
                        tma_load_2d(sb + q * 512, &tmB, bcol + kt * 16 + 4 * q, brow, &full[stage]);
                    }
                    if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    const int fr = lane >> 2, fk = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    for (int r = 0;; ++r) {
        const int t = item_of(r);
        if (t >= p.total) { if (r * G >= p.total) break; else continue; }
        const TgTile w = tg_decode(p, t);
        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        for (int kt = 0; kt < w.nk; ++kt) {
            mbar_wait(&full[stage], phase);
            const double* a_s = stage_base + (size_t)stage * (2 * SW_OPER_DOUBLES) + ((wm + fr) << 2) + fk;
            const double* b_s = stage_base + (size_t)stage * (2 * SW_OPER_DOUBLES) + SW_OPER_DOUBLES + ((wn + fr) << 2) + fk;
            double a[2][4], bb[2][8];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[0][i] = a_s[i * 32];
#pragma unroll
            for (int j = 0; j < 8; ++j) bb[0][j] = b_s[j * 32];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk < 3) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[nxt][i] = a_s[(kk + 1) * 512 + i * 32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) bb[nxt][j] = b_s[(kk + 1) * 512 + j * 32];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[cur][i], bb[cur][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
        }
        // ---- epilogue
        if (p.C) {
            double* C = p.C + p.c_off0 + (int64_t)w.z * p.c_zstep;
            if (p.beta != 0.0) {
                // read-modify-write: the eight loads of a fragment row go out together (one global-memory latency per row group
                // instead of one per element: with stores in between the compiler must keep the loads in program order)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = w.m0 + wm + i * 8 + fr;
                    double2* dst = reinterpret_cast<double2*>(C + (int64_t)row * p.ldc + w.n0 + wn + 2 * fk);
                    double2 o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = dst[j * 4];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o[j].x = fma(p.beta, o[j].x, p.alpha * acc[i][j][0]);
                        o[j].y = fma(p.beta, o[j].y, p.alpha * acc[i][j][1]);
                        dst[j * 4] = o[j];
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = w.m0 + wm + i * 8 + fr;
                    double2* dst = reinterpret_cast<double2*>(C + (int64_t)row * p.ldc + w.n0 + wn + 2 * fk);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j * 4] = make_double2(p.alpha * acc[i][j][0], p.alpha * acc[i][j][1]);
                }
            }
        }
        if (p.CT) {
            // transposed copy: a lane owns (row fr, columns 2 fk, 2 fk + 1) of an 8 x 8 fragment -> two 8-byte stores into two
            // rows of CT; the eight lanes of equal fk write 64 contiguous bytes (rows fr = 0..7 of the source)
            double* CT = p.CT + p.ct_off0 + (int64_t)w.z * p.ct_zstep;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = w.m0 + wm + i * 8 + fr;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = w.n0 + wn + j * 8 + 2 * fk;
                    CT[(int64_t)col * p.ldct + row] = p.alpha * acc[i][j][0];
                    CT[(int64_t)(col + 1) * p.ldct + row] = p.alpha * acc[i][j][1];
                }
            }
        }
    }
}

}  // namespace abo
