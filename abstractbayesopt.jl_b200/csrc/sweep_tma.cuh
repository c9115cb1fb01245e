// sweep_tma.cuh — the hot kernel of the candidate sweep, second generation:
//     partial[ib][c] = sum over the 128 rows of tile ib of ( L^-1[ib rows, 0:k_hi] * K*[c, 0:k_hi]^T )^2
// i.e. the posterior-variance contraction  ||L^-1 k*||^2  (StandardGP.jl:377-379 via AbstractGPs
// diag_Xt_invA_X) as a triangular FP64 tensor-core product with a fused sum-of-squares epilogue.
//
// Persistent, warp-specialised:
//   * one CTA per SM, static serpentine tile schedule, heaviest (largest ib) tiles first;
//   * warp 8 = TMA producer: cp.async.bulk.tensor.2d boxes of 4 (k) x 128 (rows) doubles land in
//     shared memory as [k/4][row][4] — every DMMA fragment (8 rows x 4 k) is then one 256-byte
//     contiguous, bank-conflict-free run; completion is signalled on an mbarrier ring
//     (full/empty, NSTAGE stages of 32 KB), so the producer runs ahead across tile boundaries;
//   * warps 0..7 = DMMA consumers (4 along M x 2 along N, warp tile 32 x 64), fragments
//     double-buffered in registers, no block-wide barrier inside the k loop.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm_dmma.cuh"

namespace abo {

constexpr int SW_STAGES = 5;
constexpr int SW_CONSUMERS = 256;
constexpr int SW_THREADS = SW_CONSUMERS + 32;
constexpr int SW_OPER_DOUBLES = 4 * 128 * 4;                       // one operand, one stage: 16 KB
constexpr int SW_STAGE_BYTES = 2 * SW_OPER_DOUBLES * 8;            // 32 KB
constexpr int SW_SMEM_BYTES = SW_STAGES * SW_STAGE_BYTES + 4 * 128 * 8 + 2 * SW_STAGES * 8 + 128;

struct SweepParams {
    int T;              // row tiles of L^-1 (Npad / 128)
    int ncb;            // candidate tiles in this launch
    double* sumsq;      // [T][sumsq_ld]
    int64_t sumsq_ld;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// tile t of the launch (heaviest first): ib = T-1 - t / ncb, cb = t % ncb
// CTA b processes rounds r = 0,1,...: t = r*G + (r odd ? G-1-b : b)   (serpentine: balances the
// triangular work to within one round's spread without any dynamic scheduling)
__device__ __forceinline__ int sweep_tile(int r, int b, int G) { return r * G + ((r & 1) ? (G - 1 - b) : b); }

__global__ void __launch_bounds__(SW_THREADS, 1)
sweep_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, SweepParams p) {
    extern __shared__ __align__(128) unsigned char sw_smem[];
    double* stage_base = reinterpret_cast<double*>(sw_smem);
    double* red = reinterpret_cast<double*>(sw_smem + SW_STAGES * SW_STAGE_BYTES);          // [4][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(sw_smem + SW_STAGES * SW_STAGE_BYTES + 4 * 128 * 8);
    uint64_t* empty = full + SW_STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int ntiles = p.T * p.ncb;

    if (tid == 0) {
        for (int s = 0; s < SW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int r = 0;; ++r) {
                const int t = sweep_tile(r, b, G);
                if (t >= ntiles) { if (r * G >= ntiles) break; else continue; }
                const int ib = p.T - 1 - t / p.ncb, cb = t % p.ncb;
                const int m0 = ib * 128, n0 = cb * 128;
                const int nk = (ib + 1) * 8;
                for (int kt = 0; kt < nk; ++kt) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], SW_STAGE_BYTES);
                    double* sa = stage_base + (size_t)stage * (2 * SW_OPER_DOUBLES);
                    double* sb = sa + SW_OPER_DOUBLES;
                    const int k0 = kt * 16;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        tma_load_2d(sa + q * 512, &tmA, k0 + 4 * q, m0, &full[stage]);
                        tma_load_2d(sb + q * 512, &tmB, k0 + 4 * q, n0, &full[stage]);
                    }
                    if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ------------------------------ DMMA consumers ------------------------------
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    const int fr = lane >> 2, fk = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    for (int r = 0;; ++r) {
        const int t = sweep_tile(r, b, G);
        if (t >= ntiles) { if (r * G >= ntiles) break; else continue; }
        const int ib = p.T - 1 - t / p.ncb, cb = t % p.ncb;
        const int nk = (ib + 1) * 8;

        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        for (int kt = 0; kt < nk; ++kt) {
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

        // ---- epilogue: column sums of squares over the tile's 128 rows
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) s = fma(acc[i][j][e], acc[i][j][e], s);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (fr == 0) red[(warp & 3) * 128 + wn + j * 8 + 2 * fk + e] = s;
            }
        }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");       // consumer warps only
        if (tid < 128) {
            const double s = ((red[tid] + red[128 + tid]) + red[256 + tid]) + red[384 + tid];
            p.sumsq[(int64_t)ib * p.sumsq_ld + cb * 128 + tid] = s;
        }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
    }
}


// ------------------------------------------------------------------------------------------
// Trailing update of the blocked Cholesky with the same TMA / mbarrier / DMMA pipeline:
//     C[i-tile, j-tile] -= L[i rows, kcol0 : kcol0 + 16*nk] * L[j rows, same columns]^T
// for the lower tiles (j <= i) of a rectangular range of tiles.  One tile per CTA (not
// persistent) so that the higher-priority panel stream of the look-ahead schedule can slip its
// small kernels in between tiles.  A and B are both K-contiguous row panels of the SAME
// row-major matrix, hence one tensor map.
// ------------------------------------------------------------------------------------------
// same ring depth as the sweep kernel
constexpr int SY_STAGES = 5;
constexpr int SY_SMEM_BYTES = SY_STAGES * SW_STAGE_BYTES + 4 * 128 * 8 + 2 * SY_STAGES * 8 + 128;

struct SyrkParams {
    double* C;          // base of the matrix (row-major, ld)
    int64_t ld;
    int row_t0, col_t0; // first row tile / column tile of this launch's grid
    int kcol0;          // first column of the panel block
    int nk;             // k16 steps (panel width / 16)
    int skip_first;     // leave out tile (row_t0, col_t0): it was updated by the latency kernel on the critical chain
};

__global__ void __launch_bounds__(SW_THREADS, 1)
syrk_tma_kernel(const __grid_constant__ CUtensorMap tmL, SyrkParams p) {
    extern __shared__ __align__(128) unsigned char sw_smem[];
    double* stage_base = reinterpret_cast<double*>(sw_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(sw_smem + SY_STAGES * SW_STAGE_BYTES + 4 * 128 * 8);
    uint64_t* empty = full + SY_STAGES;
    const int it = p.row_t0 + blockIdx.y, jt = p.col_t0 + blockIdx.x;
    pdl_launch_dependents();
    if (jt > it || (p.skip_first && blockIdx.x == 0 && blockIdx.y == 0)) return;
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = it * 128, n0 = jt * 128;

    if (tid == 0) {
        for (int s = 0; s < SY_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kt = 0; kt < p.nk; ++kt) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], SW_STAGE_BYTES);
                double* sa = stage_base + (size_t)stage * (2 * SW_OPER_DOUBLES);
                double* sb = sa + SW_OPER_DOUBLES;
                const int k0 = p.kcol0 + kt * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tma_load_2d(sa + q * 512, &tmL, k0 + 4 * q, m0, &full[stage]);
                    tma_load_2d(sb + q * 512, &tmL, k0 + 4 * q, n0, &full[stage]);
                }
                if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    const int fr = lane >> 2, fk = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    for (int kt = 0; kt < p.nk; ++kt) {
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
        if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + wm + i * 8 + fr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + wn + j * 8 + 2 * fk;
            double2* dst = reinterpret_cast<double2*>(p.C + (int64_t)row * p.ld + col);
            double2 o = *dst;
            o.x -= acc[i][j][0];
            o.y -= acc[i][j][1];
            *dst = o;
        }
    }
}


}  // namespace abo

// ------------------------------------------------------------------------------------------
// gemm_ws_kernel<LA, LB> — warp-specialised general tile GEMM  C = alpha * A op(B) + beta * C  for the
// operand layouts TMA boxes cannot stage conflict-free (MN-contiguous operands of the triangular
// inverse and of L^-T L^-1): two PRODUCER warps issue 16-byte cp.async copies into the same
// shared-memory layouts as gemm_dmma_kernel and signal an mbarrier ring with
// cp.async.mbarrier.arrive; eight CONSUMER warps run the DMMA inner loop of the sweep kernel with no
// block-wide barrier.  One 128 x 128 tile per CTA; k-range flags as in gemm_dmma.cuh.
// ------------------------------------------------------------------------------------------
namespace abo {

constexpr int WS_STAGES = 5;
constexpr int WS_THREADS = 256 + 64;
constexpr int WS_SMEM_BYTES = WS_STAGES * 2 * TILE_DOUBLES * (int)sizeof(double) + 2 * WS_STAGES * 8 + 64;

template <int LAYOUT>
__device__ __forceinline__ void ws_load_tile(double* s, const double* g, int64_t ld, int r0, int k0, int lane) {
    // one warp copies a 128 x 16 operand tile: 1024 chunks of 16 bytes, 32 per lane
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
        const int c = lane + 32 * q;
        if (LAYOUT == KC) {
            const int row = c >> 3, ch = c & 7;
            cp_async16(s + (((ch >> 1) * 128 + row) << 2) + ((ch & 1) << 1), g + (int64_t)(r0 + row) * ld + k0 + 2 * ch);
        } else {
            const int krow = c >> 6, ch = c & 63;
            cp_async16(s + krow * 132 + 2 * ch, g + (int64_t)(k0 + krow) * ld + r0 + 2 * ch);
        }
    }
}

// persistent tile schedule: the launch covers `total` = batch x tiles-per-matrix work items, CTA c
// runs items c, 2 grid - 1 - c, 2 grid + c, ... (serpentine).  Items are ordered heaviest first (longest k-range) with the batch
// index fastest, so a static round-robin hands every CTA one item of each weight class:
//   LOWER_ONLY (square, k >= m with KLO_M): triangular enumeration, rows ascending
//   KLO_N: column-major ascending n      KHI_M: row-major descending m      else: row-major
struct WsTile { int m0, n0, z, klo, nk; };
__device__ __forceinline__ WsTile ws_decode(const GemmParams& p, int t, int Mt, int Nt, int batch) {
    WsTile w;
    w.z = t % batch;
    const int tl = t / batch;
    int mt, nt;
    if (p.flags & LOWER_ONLY) {
        mt = (int)((sqrt(8.0 * tl + 1.0) - 1.0) * 0.5);
        while ((mt + 1) * (mt + 2) / 2 <= tl) ++mt;
        while (mt * (mt + 1) / 2 > tl) --mt;
        nt = tl - mt * (mt + 1) / 2;
    } else if (p.flags & KLO_N) { nt = tl / Mt; mt = tl % Mt; }
    else if (p.flags & KHI_M) { mt = Mt - 1 - tl / Nt; nt = tl % Nt; }
    else { mt = tl / Nt; nt = tl % Nt; }
    w.m0 = mt * BM; w.n0 = nt * BN;
    int klo = 0, khi = p.K;
    if (p.flags & KLO_M) klo = w.m0;
    if ((p.flags & KLO_N) && w.n0 > klo) klo = w.n0;
    if ((p.flags & KHI_M) && w.m0 + BM < khi) khi = w.m0 + BM;
    w.klo = klo;
    w.nk = (khi > klo) ? (khi - klo) / BK : 0;
    return w;
}

template <int LA, int LB>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_ws_kernel(GemmParams p, int Mt, int Nt, int batch, int total) {
    extern __shared__ __align__(128) unsigned char ws_smem[];
    double* sA = reinterpret_cast<double*>(ws_smem);
    double* sB = sA + WS_STAGES * TILE_DOUBLES;
    uint64_t* full = reinterpret_cast<uint64_t*>(ws_smem + WS_STAGES * 2 * TILE_DOUBLES * sizeof(double));
    uint64_t* empty = full + WS_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; ++s) { mbar_init(&full[s], 64); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;
    if (warp >= 8) {
        // ------------------------------ cp.async producers (warp 8: A, warp 9: B) ------------------------------
        for (int rnd = 0;; ++rnd) {
            const int t = rnd * (int)gridDim.x + ((rnd & 1) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x);
            if (t >= total) { if (rnd * (int)gridDim.x >= total) break; continue; }
            const WsTile w = ws_decode(p, t, Mt, Nt, batch);
            const double* A = p.A + (int64_t)w.z * p.strideA;
            const double* B = p.B + (int64_t)w.z * p.strideB;
            for (int kt = 0; kt < w.nk; ++kt) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (warp == 8) ws_load_tile<LA>(sA + stage * TILE_DOUBLES, A, p.lda, w.m0, w.klo + kt * BK, lane);
                else ws_load_tile<LB>(sB + stage * TILE_DOUBLES, B, p.ldb, w.n0, w.klo + kt * BK, lane);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&full[stage])) : "memory");
                if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");      // do not retire with copies (and their arrives) in flight
        return;
    }

    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    for (int rnd = 0;; ++rnd) {                     // serpentine: odd rounds run the CTAs in reverse order
        const int t = rnd * (int)gridDim.x + ((rnd & 1) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x);
        if (t >= total) { if (rnd * (int)gridDim.x >= total) break; continue; }
        const WsTile w = ws_decode(p, t, Mt, Nt, batch);
        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        for (int kt = 0; kt < w.nk; ++kt) {
            mbar_wait(&full[stage], phase);
            const double* a_s = sA + stage * TILE_DOUBLES;
            const double* b_s = sB + stage * TILE_DOUBLES;
            double a[2][4], bb[2][8];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[0][i] = frag<LA>(a_s, 0, wm + i * 8, lane);
#pragma unroll
            for (int j = 0; j < 8; ++j) bb[0][j] = frag<LB>(b_s, 0, wn + j * 8, lane);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk < 3) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[nxt][i] = frag<LA>(a_s, kk + 1, wm + i * 8, lane);
#pragma unroll
                    for (int j = 0; j < 8; ++j) bb[nxt][j] = frag<LB>(b_s, kk + 1, wn + j * 8, lane);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[cur][i], bb[cur][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
        }
        double* C = p.C + (int64_t)w.z * p.strideC;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = w.m0 + wm + i * 8 + (lane >> 2);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = w.n0 + wn + j * 8 + 2 * (lane & 3);
                double2* dst = reinterpret_cast<double2*>(C + (int64_t)row * p.ldc + col);
                double2 v;
                v.x = p.alpha * acc[i][j][0];
                v.y = p.alpha * acc[i][j][1];
                if (p.beta != 0.0) {
                    double2 o = *dst;
                    v.x += p.beta * o.x;
                    v.y += p.beta * o.y;
                }
                *dst = v;
            }
        }
    }
}

// sms: CTAs to launch at most (one persistent CTA per SM).  LOWER_ONLY needs a square tile grid.
template <int LA, int LB>
inline cudaError_t launch_gemm_ws(const GemmParams& p, int batch, cudaStream_t st, int sms) {
    if (p.M <= 0 || p.N <= 0 || batch <= 0) return cudaSuccess;
    const int Mt = p.M / BM, Nt = p.N / BN;
    const int64_t per = (p.flags & LOWER_ONLY) ? (int64_t)Mt * (Mt + 1) / 2 : (int64_t)Mt * Nt;
    const int64_t total = per * batch;
    if (total > 0x7fffffff) return cudaErrorInvalidValue;
    const int grid = (int)std::min<int64_t>(total, sms);
    gemm_ws_kernel<LA, LB><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(p, Mt, Nt, batch, (int)total);
    return cudaGetLastError();
}

}  // namespace abo
