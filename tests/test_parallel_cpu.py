"""Host-side multi-GPU logic on CPU: shard ranges, rank-count independent top-k merge, and a
world_size-2 gloo run of the gather/merge path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_range_partitions():
    import abo_b200 as abo
    for m in (0, 1, 7, 16_777_216, 1000003):
        for G in (1, 2, 3, 8):
            r = [abo.shard_range(m, k, G) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == m
            assert all(r[k][1] == r[k + 1][0] for k in range(G - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_merge_topk_is_rank_count_independent():
    import abo_b200 as abo
    from oracle import abo_oracle as orc
    rng = np.random.default_rng(0)
    s = rng.standard_normal(5000)
    s[rng.integers(0, 5000, 40)] = 0.25           # ties
    s[[17, 4000]] = np.nan                        # NaN first
    s[100] = -0.0; s[101] = 0.0
    ref = orc.sortperm_rev(s, 64)
    for G in (1, 2, 3, 8):
        idxs, vals = [], []
        for r in range(G):
            lo, hi = abo.shard_range(s.size, r, G)
            loc = orc.sortperm_rev(s[lo:hi], 64)
            idxs.append(loc + lo); vals.append(s[lo:hi][loc])
        gi, gv = abo.merge_topk(idxs, vals, 64)
        assert list(gi) == list(ref)
        assert np.array_equal(gv[2:], s[ref][2:]) and np.all(np.isnan(gv[:2]))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import abo_b200 as abo

    class FakeAcq:                                 # stands in for the GPU sweep on CPU
        def topk(self, surrogate, x, k):
            from oracle import abo_oracle as orc
            s = np.sin(37.0 * x[:, 0]) * np.cos(11.0 * x[:, 1])
            ti = orc.sortperm_rev(s, k)
            return s, ti, s[ti]

    rng = np.random.default_rng(3)
    cand = rng.random((10_001, 2))
    gi, gv = abo.sharded_topk(FakeAcq(), None, cand, 50)
    # NLML restarts sharded R/G (R = 7 is not divisible by 2): a closed-form stand-in for abo_nlml_batch
    theta = np.random.default_rng(5).normal(size=(7, 2))
    calls = []

    def fake_eval(th):
        calls.append(len(th))
        return (th ** 2).sum(1), 2 * th, (th[:, 0] > 1.0).astype(np.int32)

    v, g, info = abo.sharded_restarts(fake_eval, theta)
    q.put((rank, gi.tolist(), gv.tolist(), v.tolist(), g.tolist(), info.tolist(), calls))
    dist.destroy_process_group()


def test_sharded_topk_gloo_world2():
    from oracle import abo_oracle as orc
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    out = [q.get(timeout=120) for _ in range(2)]
    for p in procs: p.join(timeout=60)
    rng = np.random.default_rng(3)
    cand = rng.random((10_001, 2))
    sc = np.sin(37.0 * cand[:, 0]) * np.cos(11.0 * cand[:, 1])
    ref = orc.sortperm_rev(sc, 50)
    theta = np.random.default_rng(5).normal(size=(7, 2))
    for rank, gi, gv, v, g, info, calls in out:
        assert gi == list(ref)
        assert np.array_equal(np.array(v), (theta ** 2).sum(1)) and np.array_equal(np.array(g), 2 * theta)
        assert info == (theta[:, 0] > 1.0).astype(int).tolist()
        assert calls == [3 if rank == 0 else 4]                       # each rank evaluated only its shard
