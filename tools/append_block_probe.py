"""GradientGP block append at the C3 size (n = 512, d = 10, N = 5632): one more point (11 rows) vs a re-fit."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
c = orc.make_config("C3", m=256)
n = len(c["X"]) - 24
k = c["scale"] * abo.with_lengthscale(abo.ApproxMatern52Kernel(), 1.0 / c["inv_ls"])
gp = abo.update(abo.GradientGP(k, 11, c["noise"]), c["X"][:n], c["Y"][:n])
t0 = time.perf_counter(); abo.update(abo.GradientGP(k, 11, c["noise"]), c["X"][:n], c["Y"][:n]); t_fit = time.perf_counter() - t0
ts = []
for i in range(n, n + 24):
    t0 = time.perf_counter(); gp = abo.update(gp, c["X"][:i + 1], c["Y"][:i + 1]); ts.append(time.perf_counter() - t0)
full = abo.update(abo.GradientGP(k, 11, c["noise"]), c["X"], c["Y"], allow_append=False)
err = float(np.max(np.abs(abo.posterior_var(gp, c["Xc"]) - abo.posterior_var(full, c["Xc"]))))
print(json.dumps({"N": int(11 * (n + 24)), "refit_ms": 1e3 * t_fit, "block_append_ms_median": 1e3 * float(np.median(ts[4:])),
                  "max_abs_var_diff_vs_refit": err}))
