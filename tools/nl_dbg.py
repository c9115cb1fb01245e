import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
c = orc.make_config("C3", m=64)
k = c["scale"] * abo.with_lengthscale(abo.ApproxMatern52Kernel(), 1.0 / c["inv_ls"])
th = np.array([[np.log(1.5), 0.0], [np.log(1.0), 0.2], [np.log(2.0), -0.1], [np.log(1.2), 0.1]])
g0 = abo.GradientGP(k, 11, c["noise"])
for n in (128, 128, 128, 64, 128, 200, 200):
    t0 = time.perf_counter(); v, g, info = abo.nlml_batch(g0, th, c["X"][:n], c["Y"][:n]); t = time.perf_counter() - t0
    print("nlml R=4 n=%d N=%d ms %.2f" % (n, 11 * n, 1e3 * t), v[:2], flush=True)
gs = abo.StandardGP(abo.SqExponentialKernel(), 1e-3)
for n in (300, 300, 1000, 1000):
    t0 = time.perf_counter(); v, g, info = abo.nlml_batch(gs, th, c["X"][:n // 2], c["Y"][:n // 2, 0]); t = time.perf_counter() - t0
    print("std nlml n=%d ms %.2f" % (n // 2, 1e3 * t), flush=True)
