"""Warm timings of the full conditioning step (K build + Cholesky + triangular inverse + alpha) and of
the batched NLML, min over repetitions."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
def best(f, reps=6):
    f(); f(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)
out = {}
for n in (2048, 4096, 8192):
    c = orc.make_config("C4", n=n, m=8, d=20)
    m0 = abo.StandardGP(1.0 * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0), c["noise"])
    h = abo.GpHandle(abo.default_context(), 0, 20, 1); h.set_params(1.0, 1.0, c["noise"])
    out[f"fit_n{n}_ms"] = best(lambda: h.fit(c["X"], c["y"]))
c = orc.make_config("C5")
gp0 = abo.StandardGP(abo.SqExponentialKernel(), c["noise"])
out["nlml_256x1024_ms"] = best(lambda: abo.nlml_batch(gp0, c["theta"], c["X"], c["y"]), reps=4)
print(json.dumps(out))
