// abo_nccl.cu — multi-GPU plumbing: one process (rank) per GPU, NCCL over NVLink / NVSwitch.
// The candidate sweep and the NLML restarts shard with no data-path collective; the only
// exchanges are (i) the posterior broadcast once per BO iteration (abo_gp_sync: L, L^-1, X,
// alpha, hyper-parameters) and (ii) the all-gather of the per-rank top-k lists.
// NCCL is resolved at run time (dlopen of the copy already in the process, e.g. the one PyTorch
// bundles, else the system libnccl.so.2) so that libabo_cuda.so has no link-time dependency.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/abo.h"
#include "abo_internal.h"

namespace {
struct NcclId { char internal[128]; };
typedef int (*GetUniqueId_t)(NcclId*);
typedef int (*CommInitRank_t)(void**, int, NcclId, int);
typedef int (*CommDestroy_t)(void*);
typedef int (*Broadcast_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*AllGather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*Group_t)(void);
typedef const char* (*ErrStr_t)(int);
constexpr int NCCL_FLOAT64 = 8;

struct Nccl {
    bool ok = false;
    GetUniqueId_t getUniqueId; CommInitRank_t commInitRank; CommDestroy_t commDestroy;
    Broadcast_t broadcast; AllGather_t allGather; Group_t groupStart, groupEnd; ErrStr_t errStr;
};

Nccl& nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    void* h = nullptr;
    if (dlsym(RTLD_DEFAULT, "ncclCommInitRank")) h = RTLD_DEFAULT;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return n;
    n.getUniqueId = (GetUniqueId_t)dlsym(h, "ncclGetUniqueId");
    n.commInitRank = (CommInitRank_t)dlsym(h, "ncclCommInitRank");
    n.commDestroy = (CommDestroy_t)dlsym(h, "ncclCommDestroy");
    n.broadcast = (Broadcast_t)dlsym(h, "ncclBroadcast");
    n.allGather = (AllGather_t)dlsym(h, "ncclAllGather");
    n.groupStart = (Group_t)dlsym(h, "ncclGroupStart");
    n.groupEnd = (Group_t)dlsym(h, "ncclGroupEnd");
    n.errStr = (ErrStr_t)dlsym(h, "ncclGetErrorString");
    n.ok = n.getUniqueId && n.commInitRank && n.commDestroy && n.broadcast && n.allGather && n.groupStart && n.groupEnd;
    return n;
}
}  // namespace

#define NC(expr)                                                                                  \
    do {                                                                                          \
        int r_ = (expr);                                                                          \
        if (r_ != 0)                                                                              \
            return abo_fail(ABO_ERR_NCCL, "%s failed: %s", #expr, nccl().errStr ? nccl().errStr(r_) : "nccl error"); \
    } while (0)

void abo_nccl_teardown(abo_ctx* c) {
    if (c->nccl_comm && nccl().ok) nccl().commDestroy(c->nccl_comm);
    c->nccl_comm = nullptr;
}

extern "C" int32_t abo_nccl_unique_id(uint8_t id[128]) {
    if (!id) return abo_fail(ABO_ERR_INVALID, "null argument");
    if (!nccl().ok) return abo_fail(ABO_ERR_NCCL, "NCCL could not be loaded (libnccl.so.2)");
    NcclId u;
    NC(nccl().getUniqueId(&u));
    memcpy(id, u.internal, 128);
    return ABO_OK;
}

extern "C" int32_t abo_ctx_init_rank(abo_ctx* c, int32_t rank, int32_t nranks, const uint8_t id[128]) {
    if (!c || !id) return abo_fail(ABO_ERR_INVALID, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return abo_fail(ABO_ERR_INVALID, "bad rank %d / %d", rank, nranks);
    if (!nccl().ok) return abo_fail(ABO_ERR_NCCL, "NCCL could not be loaded (libnccl.so.2)");
    CU(cudaSetDevice(c->device));
    abo_nccl_teardown(c);
    NcclId u;
    memcpy(u.internal, id, 128);
    NC(nccl().commInitRank(&c->nccl_comm, nranks, u, rank));
    c->rank = rank; c->nranks = nranks;
    return ABO_OK;
}

// header broadcast through a small device buffer
static int bcast_doubles(abo_ctx* c, std::vector<double>& h, int root) {
    double* d;
    int rc = ws_get(c, WS_TOPK, sizeof(double) * h.size(), (void**)&d);
    if (rc) return rc;
    if (c->rank == root) CU(cudaMemcpyAsync(d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, c->stream));
    NC(nccl().broadcast(d, d, h.size(), NCCL_FLOAT64, root, c->nccl_comm, c->stream));
    CU(cudaMemcpyAsync(h.data(), d, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return ABO_OK;
}

// every rank contributes one status word; returns the first non-zero one (rank order) in *worst — the outcome of a
// multi-rank call must be COLLECTIVE: a rank that bails out alone leaves its peers blocked inside the next collective
static int gather_status(abo_ctx* c, int mine, int* worst, int* worst_rank) {
    double* d;
    int rc = ws_get(c, WS_TOPK, sizeof(double) * (size_t)(c->nranks + 1), (void**)&d);
    if (rc) return rc;
    double v = (double)mine;
    CU(cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NC(nccl().allGather(d, d + 1, 1, NCCL_FLOAT64, c->nccl_comm, c->stream));
    std::vector<double> all((size_t)c->nranks);
    CU(cudaMemcpyAsync(all.data(), d + 1, sizeof(double) * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *worst = 0; *worst_rank = -1;
    for (int r = 0; r < c->nranks; ++r)
        if ((int)all[r] != 0) { *worst = (int)all[r]; *worst_rank = r; break; }
    return ABO_OK;
}

// lower tiles (incl. the diagonal ones) of an Npad x Npad row-major matrix <-> a packed buffer of T(T+1)/2 tiles
// (tile (i, j), j <= i, at slot i(i+1)/2 + j; 128 x 128 row-major inside).  The posterior broadcast ships packed
// lower triangles: the upper tiles of L and L^-1 are never read by any kernel.
__global__ void pack_lower_tiles_kernel(const double* __restrict__ M, int64_t ld, double* __restrict__ packed, int unpack) {
    const int i = blockIdx.y, j = blockIdx.x;
    if (j > i) return;
    double* tile = packed + ((int64_t)i * (i + 1) / 2 + j) * 128 * 128;
    const double* src = M + (int64_t)i * 128 * ld + (int64_t)j * 128;
    for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
        const int r = e >> 6, c2 = e & 63;
        if (unpack) reinterpret_cast<double2*>(const_cast<double*>(src) + (int64_t)r * ld)[c2] = reinterpret_cast<const double2*>(tile + r * 128)[c2];
        else reinterpret_cast<double2*>(tile + r * 128)[c2] = reinterpret_cast<const double2*>(src + (int64_t)r * ld)[c2];
    }
}

extern "C" int32_t abo_gp_sync(abo_gp* g, int32_t root) {
    if (!g) return abo_fail(ABO_ERR_INVALID, "null gp");
    if (!g->ctx) return abo_fail(ABO_ERR_INVALID, "the context of this handle has been destroyed");
    abo_ctx* c = g->ctx;
    if (c->nranks == 1) return ABO_OK;
    if (!c->nccl_comm) return abo_fail(ABO_ERR_NCCL, "context has no NCCL communicator (abo_ctx_init_rank)");
    if (root < 0 || root >= c->nranks) return abo_fail(ABO_ERR_INVALID, "root %d out of range", root);
    CU(cudaSetDevice(c->device));
    // ---- header: ALWAYS broadcast, with the root's own status in it (h[11]) — an un-fitted root must not return
    //      before the collective its peers are already waiting in
    const int HN = 16 + 64 + 32;                       // scalars | prior means (p <= 64) | per-dimension inverse length scales (ARD)
    std::vector<double> h(HN, 0.0);
    if (c->rank == root) {
        int st = ABO_OK;
        if (!g->fitted) st = ABO_ERR_NOT_FITTED;
        else if (g->p > 64) st = ABO_ERR_INVALID;
        h[11] = (double)st;
        if (st == ABO_OK) {
            h[0] = (double)g->n; h[1] = (double)g->N; h[2] = (double)g->Npad; h[3] = (double)g->ldx; h[4] = g->kind;
            h[5] = g->d; h[6] = g->p; h[7] = g->s; h[8] = g->scale; h[9] = g->noise; h[10] = (double)g->cap_pad;
            for (int a = 0; a < g->p; ++a) h[16 + a] = g->mean_c[a];
            h[12] = g->ard ? 1.0 : 0.0;
            for (int k = 0; k < g->d && k < 32; ++k) h[80 + k] = g->sv[k];
        }
    }
    int rc = bcast_doubles(c, h, root);
    if (rc) return rc;
    if ((int)h[11] != ABO_OK)
        return abo_fail((int)h[11], (int)h[11] == ABO_ERR_NOT_FITTED ? "abo_gp_sync: the root surrogate has no posterior to broadcast"
                                                                      : "abo_gp_sync: p > 64 is not supported");
    // ---- receivers validate and allocate; the outcome is agreed on by all ranks before any bulk transfer
    int mine = ABO_OK;
    if (c->rank != root) {
        if ((int)h[5] != g->d || (int)h[6] != g->p) {
            mine = abo_fail(ABO_ERR_DIM, "abo_gp_sync: handle was created with d=%d p=%d, root has d=%d p=%d", g->d, g->p,
                            (int)h[5], (int)h[6]);
        } else {
            g->kind = (int)h[4]; g->s = h[7]; g->scale = h[8]; g->noise = h[9];
            for (int a = 0; a < g->p; ++a) g->mean_c[a] = h[16 + a];
            g->ard = h[12] != 0.0;
            g->sv.assign(g->d, g->s);
            for (int k = 0; k < g->d && k < 32; ++k) g->sv[k] = h[80 + k];
            const int64_t cap = (int64_t)h[10], ldx = (int64_t)h[3];
            g->fitted = false;
            if (cap != g->cap_pad || ldx != g->ldx || gp_shared(g)) mine = gp_alloc(g, cap, ldx);
            if (mine == ABO_OK) { g->n = (int64_t)h[0]; g->N = (int64_t)h[1]; g->Npad = (int64_t)h[2]; }
        }
    }
    const int64_t T = g->cap_pad > 0 ? g->cap_pad / 128 : (int64_t)h[10] / 128;
    const size_t packed = (size_t)(T * (T + 1) / 2) * 128 * 128;
    double* stage = nullptr;
    if (mine == ABO_OK) mine = ws_get(c, WS_KS, sizeof(double) * packed, (void**)&stage);
    std::string my_msg = mine ? abo_last_error() : "";
    int worst = 0, worst_rank = -1;
    if ((rc = gather_status(c, mine, &worst, &worst_rank))) return rc;
    if (worst != ABO_OK) {
        if (mine) return abo_fail(mine, "%s", my_msg.c_str());
        return abo_fail(worst, "abo_gp_sync: rank %d failed with status %d; no rank transferred anything", worst_rank, worst);
    }
    // ---- bulk: L and L^-1 as packed lower tiles (T(T+1)/2 of T^2: 2 x 272 MB instead of 2 x 537 MB at n = 8192),
    //      one after the other through the same staging buffer, then the vectors
    for (int which = 0; which < 2; ++which) {
        double* M = which ? g->dLinv : g->dL;
        if (c->rank == root) {
            pack_lower_tiles_kernel<<<dim3((unsigned)T, (unsigned)T), 256, 0, c->stream>>>(M, g->ld, stage, 0);
            KL(c);
        }
        NC(nccl().broadcast(stage, stage, packed, NCCL_FLOAT64, root, c->nccl_comm, c->stream));
        if (c->rank != root) {
            pack_lower_tiles_kernel<<<dim3((unsigned)T, (unsigned)T), 256, 0, c->stream>>>(M, g->ld, stage, 1);
            KL(c);
        }
    }
    NC(nccl().groupStart());
    NC(nccl().broadcast(g->dXsT, g->dXsT, (size_t)g->ldx * g->d, NCCL_FLOAT64, root, c->nccl_comm, c->stream));
    NC(nccl().broadcast(g->dAlpha, g->dAlpha, (size_t)g->cap_pad, NCCL_FLOAT64, root, c->nccl_comm, c->stream));
    NC(nccl().broadcast(g->dBeta, g->dBeta, (size_t)g->cap_pad, NCCL_FLOAT64, root, c->nccl_comm, c->stream));
    NC(nccl().broadcast(g->dDelta, g->dDelta, (size_t)g->cap_pad, NCCL_FLOAT64, root, c->nccl_comm, c->stream));
    NC(nccl().broadcast(g->dMeanC, g->dMeanC, (size_t)g->p, NCCL_FLOAT64, root, c->nccl_comm, c->stream));
    NC(nccl().groupEnd());
    CU(cudaStreamSynchronize(c->stream));
    g->fitted = true;
    return ABO_OK;
}

extern "C" int32_t abo_ctx_ranks(const abo_ctx* c, int32_t* rank, int32_t* nranks) {
    if (!c || !rank || !nranks) return abo_fail(ABO_ERR_INVALID, "null argument");
    *rank = c->rank;
    *nranks = c->nccl_comm ? c->nranks : 1;
    return ABO_OK;
}

// merge (value, index) lists with the sortperm(rev=true) order: key desc, index asc
void abo_merge_topk(std::vector<std::pair<uint64_t, int64_t>>& items, std::vector<double>& vals, int64_t k,
                    int64_t* top_idx, double* top_val, int64_t* out_count) {
    std::vector<size_t> ord(items.size());
    for (size_t i = 0; i < ord.size(); ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
        return items[a].first != items[b].first ? items[a].first > items[b].first : items[a].second < items[b].second;
    });
    const int64_t cnt = std::min<int64_t>(k, (int64_t)ord.size());
    for (int64_t i = 0; i < cnt; ++i) { top_idx[i] = items[ord[i]].second; top_val[i] = vals[ord[i]]; }
    *out_count = cnt;
}

extern "C" int32_t abo_topk_allgather(abo_ctx* c, int64_t k, int64_t count, int64_t* top_idx, double* top_val,
                                      int64_t* out_count) {
    if (!c || !top_idx || !top_val || !out_count) return abo_fail(ABO_ERR_INVALID, "null argument");
    if (k <= 0 || count < 0 || count > k) return abo_fail(ABO_ERR_INVALID, "bad k / count");
    if (c->nranks == 1) { *out_count = count; return ABO_OK; }
    if (!c->nccl_comm) return abo_fail(ABO_ERR_NCCL, "context has no NCCL communicator (abo_ctx_init_rank)");
    CU(cudaSetDevice(c->device));
    // record = [count, idx[k] (bit-cast to double), val[k]]  -> 2k+1 doubles per rank
    const size_t rec = 2 * (size_t)k + 1;
    double* d;
    int rc = ws_get(c, WS_TOPK, sizeof(double) * rec * (c->nranks + 1), (void**)&d);
    if (rc) return rc;
    std::vector<double> mine(rec, 0.0);
    mine[0] = (double)count;
    memcpy(&mine[1], top_idx, sizeof(int64_t) * count);
    memcpy(&mine[1 + k], top_val, sizeof(double) * count);
    CU(cudaMemcpyAsync(d, mine.data(), sizeof(double) * rec, cudaMemcpyHostToDevice, c->stream));
    NC(nccl().allGather(d, d + rec, rec, NCCL_FLOAT64, c->nccl_comm, c->stream));
    std::vector<double> all(rec * c->nranks);
    CU(cudaMemcpyAsync(all.data(), d + rec, sizeof(double) * rec * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<std::pair<uint64_t, int64_t>> items;
    std::vector<double> vals;
    for (int r = 0; r < c->nranks; ++r) {
        const double* q = &all[rec * r];
        const int64_t cn = (int64_t)q[0];
        for (int64_t i = 0; i < cn; ++i) {
            int64_t ix;
            memcpy(&ix, &q[1 + i], 8);
            items.emplace_back(ordkey(q[1 + k + i]), ix);
            vals.push_back(q[1 + k + i]);
        }
    }
    abo_merge_topk(items, vals, k, top_idx, top_val, out_count);
    return ABO_OK;
}

// all-gather of `count` doubles per rank (host buffers; results of the sharded NLML restarts,
// SURVEY 8e): recv holds nranks * count values in rank order
extern "C" int32_t abo_allgather_f64(abo_ctx* c, const double* send, int64_t count, double* recv) {
    if (!c || !send || !recv) return abo_fail(ABO_ERR_INVALID, "null argument");
    if (count < 0) return abo_fail(ABO_ERR_INVALID, "negative count");
    if (count == 0) return ABO_OK;
    if (c->nranks == 1) { memcpy(recv, send, sizeof(double) * count); return ABO_OK; }
    if (!c->nccl_comm) return abo_fail(ABO_ERR_NCCL, "context has no NCCL communicator (abo_ctx_init_rank)");
    CU(cudaSetDevice(c->device));
    double* d;
    int rc = ws_get(c, WS_TOPK, sizeof(double) * (size_t)count * (c->nranks + 1), (void**)&d);
    if (rc) return rc;
    CU(cudaMemcpyAsync(d, send, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    NC(nccl().allGather(d, d + count, (size_t)count, NCCL_FLOAT64, c->nccl_comm, c->stream));
    CU(cudaMemcpyAsync(recv, d + count, sizeof(double) * count * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return ABO_OK;
}
