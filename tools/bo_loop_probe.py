"""Config C1 end to end: StandardGP-SE + EI on 2-D Branin, 10 initial points, 51 loop passes, LHS grid 10 000,
100 local starts, hyper-parameters re-optimised every 10 passes.  Prints wall time and the top host costs."""
import os, sys, time, json, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc

def run(max_iter, seed=42):
    rng = np.random.default_rng(seed)
    dom = abo.ContinuousDomain([-5.0, 0.0], [10.0, 15.0])
    f = lambda x: float(orc.branin(np.asarray(x)[None, :])[0])
    X0 = dom.lower + (dom.upper - dom.lower) * rng.random((10, 2))
    y0 = [f(x) for x in X0]
    gp = abo.StandardGP(1.0 * abo.with_lengthscale(abo.SqExponentialKernel(), 3.0), 1e-6)
    bo = abo.BOStruct(f, abo.ExpectedImprovement(0.01, min(y0)), gp, dom, list(X0), y0, max_iter, 0.0)
    t0 = time.perf_counter()
    bo, acq_list, _ = abo.optimize(bo, standardize="mean_scale", hyper_params="all", num_restarts_HP=4, rng=rng)
    return time.perf_counter() - t0, bo

run(3)
dt, bo = run(50)
print(json.dumps({"c1_wall_s": dt, "passes": len(bo.xs) - 10, "best": min(float(v) for v in bo.ys_non_std),
                  "launches": abo.default_context().launch_count()}))
pr = cProfile.Profile(); pr.enable(); run(50); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
