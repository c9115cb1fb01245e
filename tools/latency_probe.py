"""Call latencies at the small end (config C1 sizes: n = 61, d = 2): what one BO iteration is made of."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
def best(f, reps=200):
    for _ in range(10): f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))
c = orc.make_config("C1", n=61, m=10_000)
k = c["scale"] * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0 / c["inv_ls"])
gp0 = abo.StandardGP(k, c["noise"])
gp = abo.update(gp0, c["X"], c["y"])
acq = abo.ExpectedImprovement(0.01, float(np.min(c["y"])))
th = np.array([[0.0, 0.0], [0.5, 0.2], [-0.3, 0.1], [1.0, -0.5]])
out = {
    "fit_n61_us": best(lambda: abo.update(gp0, c["X"], c["y"], allow_append=False)),
    "append_n60_us": best(lambda: abo.update(abo.update(gp0, c["X"][:60], c["y"][:60]), c["X"], c["y"]), 50),
    "sweep_topk_10k_us": best(lambda: acq.topk(gp, c["Xc"], 100)),
    "acq_value_grad_100_us": best(lambda: acq.value_and_grad(gp, c["Xc"][:100])),
    "posterior_1pt_us": best(lambda: abo.posterior_mean(gp, c["Xc"][:1])),
    "nlml_batch_R4_us": best(lambda: abo.nlml_batch(gp0, th, c["X"], c["y"])),
}
print(json.dumps(out))
