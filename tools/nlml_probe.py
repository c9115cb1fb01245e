"""C5: 256 restarts of NLML + gradient at n = 1024, d = 8 (used under ncu for launch lists)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
c = orc.make_config("C5")
gp0 = abo.StandardGP(abo.SqExponentialKernel(), c["noise"])
for it in range(3):
    t0 = time.perf_counter(); v, g, info = abo.nlml_batch(gp0, c["theta"], c["X"], c["y"]); dt = time.perf_counter() - t0
print(f"nlml_batch R=256 n=1024: {dt*1e3:.2f} ms  ({256*1024.0**3/dt/1e12:.2f} TFLOP/s)  ok={int((info==0).sum())}")
