import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, abo_b200 as abo
from oracle import abo_oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
c = orc.make_config("C4", n=n, m=8, d=20)
h = abo.GpHandle(abo.default_context(), 0, 20, 1); h.set_params(1.0, 1.0, c["noise"])
h.fit(c["X"], c["y"]); h.fit(c["X"], c["y"])
print("ok")
