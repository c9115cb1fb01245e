"""Sizes beyond the headline: fit + sweep at n = 16384 (checked against the CPU restatement on a sample) and
n = 32768 (self-consistency: interpolation at the data, variance bounds), with timings."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abo_b200 as abo
from oracle import abo_oracle as orc
for n in [int(a) for a in sys.argv[1:]] or [16384]:
    c = orc.make_config("C4", n=n, m=20000, d=20)
    k = c["scale"] * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0 / c["inv_ls"])
    t0 = time.perf_counter(); gp = abo.update(abo.StandardGP(k, c["noise"]), c["X"], c["y"]); t_fit = time.perf_counter() - t0
    t0 = time.perf_counter(); gp = abo.update(abo.StandardGP(k, c["noise"]), c["X"], c["y"]); t_fit2 = time.perf_counter() - t0
    acq = abo.UpperConfidenceBound(2.0)
    t0 = time.perf_counter(); s, ti, tv = acq.topk(gp, c["Xc"], 100); t_sw = time.perf_counter() - t0
    res = {"n": n, "fit_ms_first": 1e3 * t_fit, "fit_ms": 1e3 * t_fit2, "fit_tflops": (n ** 3 / 3 * 2) / t_fit2 / 1e12,
           "sweep_20k_ms": 1e3 * t_sw, "sweep_tflops": 20000 * float(n) ** 2 / t_sw / 1e12}
    mu_d = abo.posterior_mean(gp, c["X"][:512]); var_d = abo.posterior_var(gp, c["X"][:512])
    res["max_resid_at_data"] = float(np.max(np.abs(mu_d - c["y"][:512])))
    res["var_at_data_range"] = [float(var_d.min()), float(var_d.max())]
    assert np.all(var_d > 0) and np.all(var_d < c["noise"] * 1.01), "posterior variance at the data must be in (0, noise)"
    assert np.all(np.isfinite(s)) and np.array_equal(tv, s[ti])
    if n <= 16384:
        t0 = time.perf_counter(); post = orc.fit_standard(c["X"], c["y"], c["kind"], c["inv_ls"], c["scale"], c["noise"]); res["oracle_fit_s"] = time.perf_counter() - t0
        sel = np.concatenate([ti[:20], np.arange(0, 20000, 97)])
        mu_o, var_o = orc.posterior_mean_var(post, c["Xc"][sel])
        ref = orc.acquisition(2, acq.params(), mu_o, var_o)
        res["max_rel_err_scores"] = float(np.max(np.abs(s[sel] - ref)) / np.max(np.abs(ref)))
        assert res["max_rel_err_scores"] < 1e-9
    print(json.dumps(res), flush=True)
    del gp
