"""The README's tutorial-style example, executed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, abo_b200 as abo

f = lambda x: (x[0]**2 + x[1] - 11)**2 + (x[0] + x[1]**2 - 7)**2          # Himmelblau
domain = abo.ContinuousDomain([-6.0, -6.0], [6.0, 6.0])
rng = np.random.default_rng(0)
xs = [domain.lower + (domain.upper - domain.lower) * rng.random(2) for _ in range(5)]
ys = [f(x) for x in xs]
model = abo.StandardGP(1.0 * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0), 1e-9)
acq = abo.ExpectedImprovement(0.0, min(ys))
bo = abo.BOStruct(f, acq, model, domain, xs, ys, 50, 0.0)
t0 = time.perf_counter()
bo, acq_values, (mu, sigma) = abo.optimize(bo, standardize="mean_only", hyper_params="all", num_restarts_HP=4, rng=rng)
print("wall s", time.perf_counter() - t0, "points", len(bo.xs), "stopped early", bo.flag)
print(min(bo.ys_non_std), bo.xs[int(np.argmin(bo.ys_non_std))])
