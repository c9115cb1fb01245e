"""Warm fit / sweep timings over n (handle re-used, 3 repetitions, best)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
for n in [int(a) for a in sys.argv[1:]]:
    c = orc.make_config("C4", n=n, m=16384, d=20)
    h = abo.GpHandle(abo.default_context(), 0, 20, 1); h.set_params(1.0, 1.0, c["noise"])
    tf = []
    for _ in range(3):
        t0 = time.perf_counter(); h.fit(c["X"], c["y"]); tf.append(time.perf_counter() - t0)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); h.posterior(c["Xc"]); ts.append(time.perf_counter() - t0)
    print(json.dumps({"n": n, "fit_ms": [round(1e3 * t, 2) for t in tf], "fit_tflops": (2 * n ** 3 / 3) / min(tf) / 1e12,
                      "sweep16k_ms": [round(1e3 * t, 2) for t in ts], "sweep_tflops": 16384 * float(n) ** 2 / min(ts) / 1e12}), flush=True)
    h.close()
