"""Generates tests/golden/known_answers.json: the closed-form known answers of the reference's own
test-suite for this path, evaluated with 50-digit mpmath (independent of oracle/ and of the CUDA
path).  The reference is Julia and cannot be imported here; these are the formulas written inline in
/root/reference/test/test_surrogates.jl:59-105,145-170,235-291 and test_acquisition.jl:20-43 (inputs
copied from those tests), plus EI/PI/UCB evaluated from the source formulas
(ExpectedImprovement.jl:59-66, ProbabilityImprovement.jl:57-63, UpperConfidenceBound.jl:38-45).
Run:  python tests/golden/make_known_answers.py"""
import json
import os

import mpmath as mp

mp.mp.dps = 50


def se(a, b):
    return mp.e ** (-(sum((mp.mpf(x) - mp.mpf(y)) ** 2 for x, y in zip(a, b))) / 2)


def posterior_1d(xs, ys, noise, xstar):
    n = len(xs)
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = se([xs[i]], [xs[j]]) + (mp.mpf(noise) if i == j else 0)
    k = mp.matrix([se([xstar], [x]) for x in xs])
    y = mp.matrix([mp.mpf(v) for v in ys])
    mean = (k.T * mp.lu_solve(K, y))[0]
    var = 1 - (k.T * mp.lu_solve(K, k))[0] + mp.mpf("1e-18")
    nlml = (y.T * mp.lu_solve(K, y))[0] / 2 + mp.log(mp.det(K)) / 2 + n * mp.log(2 * mp.pi) / 2
    return mean, var, nlml


out = {}
m, v, nl = posterior_1d(["0.0", "0.5", "1.0"], ["0.0", "0.25", "1.0"], "0.1", "0.25")
out["G1"] = {"source": "test/test_surrogates.jl:59-105", "mean": float(m), "var": float(v)}
out["G2"] = {"source": "test/test_surrogates.jl:145-170", "nlml": float(nl)}
m, v, _ = posterior_1d(["0.0", "0.5", "1.0"], ["2.0", "1.0", "0.5"], "0.1", "0.25")
delta = mp.mpf("0.5") - mp.mpf("0.01") - m
s = mp.sqrt(v); z = delta / s
out["G3"] = {"source": "test/test_acquisition.jl:20-43,74-95,126-148 + acquisition source formulas",
             "mean": float(m), "var": float(v), "EI": float(delta * mp.ncdf(z) + s * mp.npdf(z)),
             "PI": float(mp.ncdf(z)), "UCB_beta2": float(-m + 2 * s)}
k = mp.e ** mp.mpf("-0.01")
out["G4"] = {"source": "test/test_surrogates.jl:235-291 (SE, x=[.5,.5], y=[.6,.6])", "k": float(k),
             "dk_dy": float(-k / 10), "dk_dx": float(k / 10), "d2k_diag": float(k * (1 - mp.mpf("0.01"))),
             "d2k_offdiag": float(-k * mp.mpf("0.01"))}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
