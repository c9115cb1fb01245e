"""One GradientGP conditioning step at the C3 size (n = 512, d = 10, N = 5632) and one small batched GradientGP NLML;
used under ncu for launch lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
c = orc.make_config("C3", m=64)
k = c["scale"] * abo.with_lengthscale(abo.ApproxMatern52Kernel(), 1.0 / c["inv_ls"])
for _ in range(2):
    t0 = time.perf_counter(); gp = abo.update(abo.GradientGP(k, 11, c["noise"]), c["X"], c["Y"]); t = time.perf_counter() - t0
print("fit ms", 1e3 * t)
th = np.array([[np.log(1.5), 0.0], [np.log(1.0), 0.2], [np.log(2.0), -0.1], [np.log(1.2), 0.1]])
g0 = abo.GradientGP(k, 11, c["noise"])
for _ in range(4):
    t0 = time.perf_counter(); v, g, info = abo.nlml_batch(g0, th, c["X"][:128], c["Y"][:128]); t = time.perf_counter() - t0
    print("nlml R=4 N=1408 ms", 1e3 * t, info, flush=True)
