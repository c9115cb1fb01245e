"""Determinism stress of the sharded BO iteration (bench.py's step) at N ranks:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_stress.py [--nobs 8192] [--cands M] [--iters K]
Every iteration repeats clone + append + posterior broadcast + sweep + top-k all-gather on IDENTICAL inputs; each rank
compares the bit pattern of its whole score vector and of its local / the global top-100 with iteration 0 and reports
every difference (count, the candidate tiles they fall in, largest deviation).  Exit code 1 on any difference."""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import abo_b200 as abo  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nobs", type=int, default=8192)
ap.add_argument("--d", type=int, default=20)
ap.add_argument("--cands", type=int, default=1 << 18)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--topk", type=int, default=100)
args = ap.parse_args()
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
ctx = abo.Context(lrank)
if world > 1:
    abo.init_nccl_context(ctx)
n, d, m = args.nobs, args.d, args.cands
rng = np.random.default_rng(7)
X = rng.random((n, d)); y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n)
kern = 1.0 * abo.with_lengthscale(abo.SqExponentialKernel(), 1.5)
base = abo.update(abo.StandardGP(kern, 1e-6, ctx=ctx), X[:-1], y[:-1]) if rank == 0 else None
recv = abo.empty_posterior_like(abo.StandardGP(kern, 1e-6, ctx=ctx), d) if rank != 0 else None
acq = abo.ExpectedImprovement(0.01, float(y.min()))
gen = torch.Generator(device="cuda"); gen.manual_seed(1234 + rank)
Xc = torch.rand((m, d), dtype=torch.float64, device="cuda", generator=gen)
scores = torch.empty(m, dtype=torch.float64, device="cuda")
ref = None
bad = []
for it in range(args.iters):
    if rank == 0:
        h = base.gpx.clone(); h.append(X[-1], y[-1:])
    else:
        h = recv.gpx
    if world > 1:
        h.sync(0)
    scores.fill_(float("nan"))
    torch.cuda.synchronize()
    ti, tv = h.acq_eval_dev(acq.acq_id, acq.params(), Xc.data_ptr(), m, scores.data_ptr(), k=args.topk)
    gi, gv = (ctx.topk_allgather(args.topk, ti + rank * m, tv) if world > 1 else (ti, tv))
    torch.cuda.synchronize()
    cur = (scores.clone(), np.array(ti), np.array(tv), np.array(gi), np.array(gv))
    if rank == 0:
        h.close()
    if ref is None:
        ref = cur
        continue
    diff = (cur[0].view(torch.int64) != ref[0].view(torch.int64))
    nd = int(diff.sum().item())
    same_local = np.array_equal(cur[1], ref[1]) and np.array_equal(cur[2].view(np.int64), ref[2].view(np.int64))
    same_glob = np.array_equal(cur[3], ref[3]) and np.array_equal(cur[4].view(np.int64), ref[4].view(np.int64))
    if nd or not same_local or not same_glob:
        ix = torch.nonzero(diff).flatten()[:4096].cpu().numpy()
        dev = float((cur[0] - ref[0]).abs().nan_to_num(nan=float("inf")).max().item()) if nd else 0.0
        bad.append({"iter": it, "rank": rank, "scores_differing": nd, "max_abs_dev": dev, "tiles": sorted(set((ix // 128).tolist()))[:32],
                    "first_idx": ix[:16].tolist(), "local_topk_same": bool(same_local), "global_topk_same": bool(same_glob)})
allbad = [None] * world
if world > 1:
    dist.all_gather_object(allbad, bad)
else:
    allbad = [bad]
if rank == 0:
    flat = [b for per in allbad for b in per]
    print(json.dumps({"ranks": world, "n": n, "cands_per_rank": m, "iters": args.iters, "deterministic": not flat, "differences": flat[:40]}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(1 if any(allbad) and any(len(b) for b in allbad) else 0)
