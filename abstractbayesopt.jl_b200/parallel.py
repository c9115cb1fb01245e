"""Multi-GPU host logic: one process per GPU.  Candidates (and NLML restarts) are sharded by
contiguous index range with no data-path collective; the posterior is broadcast once per BO
iteration with NCCL (abo_gp_sync) and the per-rank top-k lists are merged with the
(value desc, index asc, NaN first) order of sortperm(scores; rev=true) (acq_utils.jl:51) so the
selected candidates do not depend on the number of ranks."""
from __future__ import annotations

import numpy as np


def shard_range(m: int, rank: int, world: int):
    """Rank r owns candidates [r*m/G, (r+1)*m/G)."""
    return (rank * m) // world, ((rank + 1) * m) // world


def _ordkey(v: np.ndarray) -> np.ndarray:
    """Julia isless order on Float64 as uint64 keys (NaN largest, -0.0 < 0.0)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    u = v.view(np.uint64)
    neg = (u >> np.uint64(63)).astype(bool)
    key = np.where(neg, ~u, u | np.uint64(1 << 63))
    return np.where(np.isnan(v), np.uint64(0xFFFFFFFFFFFFFFFF), key)


def merge_topk(idx_lists, val_lists, k: int):
    """Merge per-rank (global index, value) lists into the global stable descending top-k."""
    idx = np.concatenate([np.asarray(a, dtype=np.int64) for a in idx_lists]) if idx_lists else np.empty(0, np.int64)
    val = np.concatenate([np.asarray(a, dtype=np.float64) for a in val_lists]) if val_lists else np.empty(0)
    if idx.size == 0:
        return idx, val
    key = _ordkey(val)
    order = np.lexsort((idx, np.iinfo(np.uint64).max - key))
    order = order[:min(k, order.size)]
    return idx[order], val[order]


def sharded_topk(acq, surrogate, candidates, k: int, group=None):
    """Every rank holds the full candidate array (or can index its shard), evaluates its shard
    with the fused sweep, and all ranks end up with the same global top-k (global indices)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(len(candidates), rank, world)
    _, ti, tv = acq.topk(surrogate, candidates[lo:hi], k)
    mine = (np.asarray(ti, dtype=np.int64) + lo, np.asarray(tv))
    if world == 1:
        return merge_topk([mine[0]], [mine[1]], k)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    return merge_topk([g[0] for g in gathered], [g[1] for g in gathered], k)


def init_nccl_context(ctx, group=None):
    """Give `ctx` an NCCL rank matching the torch.distributed group: rank 0 creates the NCCL
    unique id, it is broadcast through the existing process group (any backend)."""
    import torch.distributed as dist
    from ._lib import nccl_unique_id
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.init_rank(rank, world, box[0])


def sync_posterior(model, root: int = 0):
    """Broadcast the conditioned surrogate (L, L^-1, X, alpha, hyper-parameters) from `root`
    to every rank over NCCL / NVLink.  Non-root models must already hold a handle of the same
    (kernel, d, p) — e.g. created by update() on any data, or by `empty_like`."""
    model.gpx.sync(root)
    return model
