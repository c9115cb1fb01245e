"""Wall-clock (device-synchronised) timings of the BASELINE.json configurations through the public
host API: fit, fused sweep, NLML batch.  Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc

KN = {0: "SqExponentialKernel", 1: "Matern52Kernel", 3: "ApproxMatern52Kernel"}
def kern(c): return c["scale"] * abo.with_lengthscale(abo.Kernel(KN[c["kind"]]), 1.0 / c["inv_ls"])
def timed(f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t0)
    return best, r

out = {}
# C2: Matern-5/2, d = 6, n = 2048, EI over 1M candidates
c = orc.make_config("C2")
t_fit, gp = timed(lambda: abo.update(abo.StandardGP(kern(c), c["noise"]), c["X"], c["y"]))
acq = abo.ExpectedImprovement(*c["acq_params"])
t_sw, _ = timed(lambda: acq.topk(gp, c["Xc"], 100))
fl = (2048.0 ** 2 + 3 * 2048 * 6 + 4 * 2048) * len(c["Xc"])
out["C2"] = {"fit_ms": 1e3 * t_fit, "sweep_1M_ms": 1e3 * t_sw, "cand_per_s": len(c["Xc"]) / t_sw, "tflops_incl_copies": fl / t_sw / 1e12}
# C3: GradientGP n = 512, d = 10 -> N = 5632, sweep 65536
c = orc.make_config("C3")
t_fit, gp = timed(lambda: abo.update(abo.GradientGP(kern(c), 11, c["noise"]), c["X"], c["Y"]))
acq = abo.ExpectedImprovement(*c["acq_params"])
t_sw, _ = timed(lambda: acq.topk(gp, c["Xc"], 100))
out["C3"] = {"fit_ms": 1e3 * t_fit, "sweep_64k_ms": 1e3 * t_sw, "cand_per_s": len(c["Xc"]) / t_sw,
             "tflops": 5632.0 ** 2 * len(c["Xc"]) / t_sw / 1e12}
# C4: fit at n = 8192 and an appended observation
c = orc.make_config("C4", m=8)
t_fit, gp = timed(lambda: abo.update(abo.StandardGP(kern(c), c["noise"]), c["X"][:-1], c["y"][:-1]), reps=2)
t_app, _ = timed(lambda: abo.update(gp, c["X"], c["y"]), reps=3)
out["C4"] = {"fit_8191_ms": 1e3 * t_fit, "append_incl_clone_ms": 1e3 * t_app}
# C5: 256 restarts NLML + gradient, n = 1024, d = 8
c = orc.make_config("C5")
gp0 = abo.StandardGP(abo.SqExponentialKernel(), c["noise"])
t_nl, _ = timed(lambda: abo.nlml_batch(gp0, c["theta"], c["X"], c["y"]))
out["C5"] = {"nlml_grad_256_ms": 1e3 * t_nl, "tflops": 256 * 1024.0 ** 3 / t_nl / 1e12}
# C1: one BO iteration at tutorial size (n = 30): grid sweep + batched refinement
c = orc.make_config("C1", n=30, m=10)
gp = abo.update(abo.StandardGP(kern(c), 1e-6), c["X"], c["y"])
acq = abo.ExpectedImprovement(0.01, float(c["y"].min()))
dom = abo.ContinuousDomain(c["lower"], c["upper"])
t_oa, _ = timed(lambda: abo.optimize_acquisition(acq, gp, dom, rng=np.random.default_rng(0)))
t_oas, _ = timed(lambda: abo.optimize_acquisition(acq, gp, dom, rng=np.random.default_rng(0), refine="scipy"), reps=1)
out["C1"] = {"optimize_acquisition_batched_ms": 1e3 * t_oa, "optimize_acquisition_sequential_fd_ms": 1e3 * t_oas}
print(json.dumps(out, indent=1))
