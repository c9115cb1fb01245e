"""Sweep throughput away from the headline size: n in {64, 256, 512, 1024, 2048}, d = 6, Matern-5/2, EI + top-100 over
1 M host candidates (copies included) with the per-kernel device shares."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
out = []
ctx = abo.default_context()
for n in (64, 256, 512, 1024, 2048):
    c = orc.make_config("C2", n=n, m=1 << 20)
    k = c["scale"] * abo.with_lengthscale(abo.Matern52Kernel(), 1.0 / c["inv_ls"])
    gp = abo.update(abo.StandardGP(k, c["noise"]), c["X"], c["y"])
    acq = abo.ExpectedImprovement(0.01, float(np.min(c["y"])))
    acq.topk(gp, c["Xc"], 100)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); acq.topk(gp, c["Xc"], 100); ts.append(time.perf_counter() - t0)
    ctx.profile(True); acq.topk(gp, c["Xc"], 100); ms, cnt = ctx.profile_read(); ctx.profile(False)
    t = min(ts)
    out.append({"n": n, "ms": 1e3 * t, "cand_per_s": len(c["Xc"]) / t, "ks_build_ms": ms[0], "contraction_ms": ms[1],
                "epilogue_ms": ms[2], "launches": cnt[0]})
    print(json.dumps(out[-1]), flush=True)
