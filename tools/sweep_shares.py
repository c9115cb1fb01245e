"""Per-kernel shares of the fused sweep for a given config (C2 / C3 / C4) via abo_ctx_profile."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
KN = {0: "SqExponentialKernel", 1: "Matern52Kernel", 3: "ApproxMatern52Kernel"}
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
kw = {"C2": {}, "C3": {}, "C4": dict(m=262144)}[name]
c = orc.make_config(name, **kw)
k = c["scale"] * abo.with_lengthscale(abo.Kernel(KN[c["kind"]]), 1.0 / c["inv_ls"])
gp = abo.GradientGP(k, c["X"].shape[1] + 1, c["noise"]) if name == "C3" else abo.StandardGP(k, c["noise"])
gp = abo.update(gp, c["X"], c["Y"] if name == "C3" else c["y"])
acq = abo.ExpectedImprovement(0.01, 0.0)
ctx = abo.default_context()
acq.topk(gp, c["Xc"], 100)
t0 = time.perf_counter(); acq.topk(gp, c["Xc"], 100); t_all = time.perf_counter() - t0
t0 = time.perf_counter(); acq(gp, c["Xc"]); t_nok = time.perf_counter() - t0
ctx.profile(True); acq.topk(gp, c["Xc"], 100); ms, n = ctx.profile_read(); ctx.profile(False)
print(json.dumps({"config": name, "m": len(c["Xc"]), "call_ms_with_topk": 1e3 * t_all, "call_ms_scores_only": 1e3 * t_nok,
                  "device_ms": {"ks_build": ms[0], "contraction": ms[1], "epilogue": ms[2]}, "launches": n}))
