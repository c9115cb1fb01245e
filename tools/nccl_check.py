"""2+ rank check of the NCCL paths (run under torchrun on a multi-GPU box):
posterior broadcast (abo_gp_sync) gives bit-identical predictions on every rank, and the
sharded sweep + abo_topk_allgather selects exactly the single-GPU top-k."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import abo_b200 as abo
from oracle import abo_oracle as orc

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = abo.default_context(lr)
abo.init_nccl_context(ctx)
c = orc.make_config("C4", n=1500, m=40_000, d=20)
model = abo.StandardGP(c["scale"] * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0 / c["inv_ls"]), c["noise"], ctx=ctx)
if rank == 0:
    model = abo.update(model, c["X"], c["y"])
else:
    model = abo.empty_posterior_like(model, 20)
import time
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
abo.sync_posterior(model, 0)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
acq = abo.UpperConfidenceBound(2.0)
mu = abo.posterior_mean(model, c["Xc"][:2000]); var = abo.posterior_var(model, c["Xc"][:2000])
t = torch.from_numpy(np.concatenate([mu, var])).cuda()
ref = t.clone(); dist.broadcast(ref, 0)
assert torch.equal(t, ref), "posterior differs between ranks after abo_gp_sync"
lo, hi = abo.shard_range(len(c["Xc"]), rank, world)
_, ti, tv = acq.topk(model, c["Xc"][lo:hi], 100)
gi, gv = ctx.topk_allgather(100, ti + lo, tv)
gi2, gv2 = abo.sharded_topk(acq, model, c["Xc"], 100)
assert list(gi) == list(gi2) and np.array_equal(gv, gv2)
# NLML restarts sharded R/G per rank, results all-gathered over NCCL: identical to the single-GPU batch
c5 = orc.make_config("C5", n=300, m=13)
gp0 = abo.StandardGP(abo.SqExponentialKernel(), c5["noise"], ctx=ctx)
v_s, g_s, i_s = abo.sharded_nlml_batch(gp0, c5["theta"], c5["X"], c5["y"])
v_1, g_1, i_1 = abo.nlml_batch(gp0, c5["theta"], c5["X"], c5["y"])
assert np.array_equal(v_s, v_1) and np.array_equal(g_s, g_1) and np.array_equal(i_s, i_1), "sharded NLML differs"
# GradientGP: broadcast, then every rank appends the same new point (block append on the received posterior)
c3 = orc.make_config("C3", n=40, m=300, d=4)
kg = c3["scale"] * abo.with_lengthscale(abo.ApproxMatern52Kernel(), 1.0 / c3["inv_ls"])
gm = abo.GradientGP(kg, 5, c3["noise"], ctx=ctx)
gm = abo.update(gm, c3["X"][:39], c3["Y"][:39]) if rank == 0 else abo.empty_posterior_like(gm, 4)
abo.sync_posterior(gm, 0)
gm.gpx.append(c3["X"][39], c3["Y"][39])
mu_g = abo.posterior_grad_mean(gm, c3["Xc"][:100])
tg = torch.from_numpy(mu_g).cuda(); rg = tg.clone(); dist.broadcast(rg, 0)
assert torch.equal(tg, rg), "GradientGP posterior differs between ranks after sync + block append"
full = abo.update(abo.GradientGP(kg, 5, c3["noise"], ctx=ctx), c3["X"], c3["Y"])
assert np.max(np.abs(mu_g - abo.posterior_grad_mean(full, c3["Xc"][:100]))) < 1e-9
# the outcome of abo_gp_sync is collective: an un-fitted root, or a receiver created with another dimension, makes EVERY
# rank return an error instead of leaving its peers blocked inside the broadcast
unf = abo.empty_posterior_like(abo.StandardGP(abo.SqExponentialKernel(), 0.1, ctx=ctx), 20)
try:
    unf.gpx.sync(0); raised = False
except abo.AboCudaError as e:
    raised = "no posterior" in str(e)
assert raised, "un-fitted root must fail on every rank"
wrong = abo.empty_posterior_like(abo.StandardGP(abo.SqExponentialKernel(), 0.1, ctx=ctx), 20 if rank == 0 else 7)
src = model.gpx if rank == 0 else wrong.gpx
try:
    src.sync(0); raised = False
except (abo.DimensionMismatch, abo.AboCudaError) as e:
    raised = True
assert raised, "a receiver of the wrong dimension must fail the sync on every rank"
abo.sync_posterior(model, 0)                                  # and the communicator is still usable afterwards
if rank == 0:
    s_all, ti_all, tv_all = acq.topk(model, c["Xc"], 100)
    assert list(ti_all) == list(gi), "sharded top-k differs from the single-GPU top-k"
    N = 1536
    print(f"nccl_check ok: world={world} sync of {2 * N * N * 8 / 1e6:.1f} MB in {dt * 1e3:.2f} ms; top-k, sharded NLML and GradientGP sync+append identical", flush=True)
dist.destroy_process_group()
