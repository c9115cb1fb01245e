"""One warm call of the batched NLML + gradient on config C5 (256 restarts, n = 1024, d = 8); run under
`ncu --metrics gpu__time_duration.sum` for the launch list (profiles/launches_nlml_c5_r02.csv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import abo_b200 as abo
from oracle import abo_oracle as orc
c = orc.make_config("C5")
gp0 = abo.StandardGP(abo.SqExponentialKernel(), c["noise"])
for _ in range(3):
    abo.nlml_batch(gp0, c["theta"], c["X"], c["y"])
