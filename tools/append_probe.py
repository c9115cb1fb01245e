"""O(n^2) row append at n = 8192: wall time of abo_gp_append alone and its HBM roofline fraction
(algorithmic bytes: two passes over the packed lower triangle of L^-1)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
c = orc.make_config("C4", n=n + 40, m=8, d=20)
gp = abo.update(abo.StandardGP(c["scale"] * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0), c["noise"]), c["X"][:n], c["y"][:n])
h = gp.gpx.clone()
ts = []
for i in range(n, n + 40):
    t0 = time.perf_counter(); h.append(c["X"][i], c["y"][i:i + 1]); ts.append(time.perf_counter() - t0)
t = float(np.median(ts[5:]))
t0 = time.perf_counter(); h2 = gp.gpx.clone(); t_clone = time.perf_counter() - t0
# the BO loop's step (bayesian_opt.jl:113-150): snapshot copy + functional update with one more observation;
# the update un-shares the snapshot's buffers (copy-on-write) and appends
tb = []
g = gp
for i in range(n, n + 12):
    t0 = time.perf_counter(); prev = g.copy(); g = abo.update(g, c["X"][:i + 1], c["y"][:i + 1]); tb.append(time.perf_counter() - t0)
t_bo = float(np.median(tb[3:]))
bytes_alg = 2 * (n * (n + 1) / 2) * 8
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6548.2
print(json.dumps({"n": n, "append_ms_median": 1e3 * t, "algorithmic_GB": bytes_alg / 1e9, "achieved_GBs": bytes_alg / t / 1e9,
                  "hbm_peak_GBs": peak, "frac": bytes_alg / t / 1e9 / peak, "clone_ms": 1e3 * t_clone,
                  "bo_update_ms_copy_plus_append": 1e3 * t_bo}))
