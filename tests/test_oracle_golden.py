"""Pins oracle/abo_oracle.py against the reference's own known-answer tests (closed forms
written inline in /root/reference/test/*.jl) and against a 50-digit mpmath arbiter.
Each test names the reference test it reproduces."""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import abo_oracle as orc

mp.mp.dps = 50


def se(a, b, s=1.0, sc=1.0):
    a, b = np.atleast_1d(a), np.atleast_1d(b)
    return sc * math.exp(-0.5 * float(np.sum((s * a - s * b) ** 2)))


# ---- test/test_surrogates.jl:59-105  (G1)
def test_standard_gp_posterior_known_answer():
    xs = np.array([[0.0], [0.5], [1.0]]); ys = np.array([0.0, 0.25, 1.0]); noise = 0.1
    post = orc.fit_standard(xs, ys, orc.SE, 1.0, 1.0, noise)
    mean, var = orc.posterior_mean_var(post, np.array([[0.25]]))
    K = np.array([[se(a, b) for b in xs] for a in xs]) + noise * np.eye(3)
    k = np.array([se([0.25], a) for a in xs])
    true_mean = k @ np.linalg.solve(K, ys)
    true_var = se([0.25], [0.25]) - k @ np.linalg.solve(K, k)
    assert abs(mean[0] - true_mean) < 1e-10 and abs(var[0] - true_var) < 1e-10
    # SURVEY §8c golden values (cross-checked there with 50-digit mpmath)
    assert abs(mean[0] - 0.1771247751991296) < 1e-14
    assert abs(var[0] - 0.050320225208722924) < 1e-14
    mm, vv = orc.mp_posterior_standard(xs, ys, orc.SE, 1.0, 1.0, noise, 0.0, np.array([[0.25]]))
    assert abs(float(mm[0]) - mean[0]) < 1e-15 and abs(float(vv[0]) - var[0]) < 1e-15


# ---- test/test_surrogates.jl:145-170  (G2)
def test_standard_gp_nlml_known_answer():
    xs = np.array([[0.0], [0.5], [1.0]]); ys = np.array([0.0, 0.25, 1.0]); noise = 0.1
    val = orc.nlml(xs, ys, orc.SE, math.log(1.0), math.log(1.0), noise)
    K = np.array([[se(a, b) for b in xs] for a in xs]) + noise * np.eye(3)
    true = 0.5 * (ys @ np.linalg.solve(K, ys) + np.linalg.slogdet(K)[1] + 3 * math.log(2 * math.pi))
    assert abs(val - true) < 1e-10
    assert abs(val - 2.6769327097262567) < 1e-13


# ---- test/test_acquisition.jl:20-43,74-95,126-148 + source formulas (G3)
def test_acquisition_known_answers():
    xs = np.array([[0.0], [0.5], [1.0]]); ys = np.array([2.0, 1.0, 0.5]); noise = 0.1
    post = orc.fit_standard(xs, ys, orc.SE, 1.0, 1.0, noise)
    mu, var = orc.posterior_mean_var(post, np.array([[0.25]]))
    assert abs(mu[0] - 1.467255970550952) < 1e-13
    assert abs(var[0] - 0.050320225208722924) < 1e-14
    ei = orc.expected_improvement(mu, var, 0.01, 0.5)[0]
    pi = orc.probability_improvement(mu, var, 0.01, 0.5)[0]
    ucb = orc.upper_confidence_bound(mu, var, 2.0)[0]
    # independent mpmath evaluation of ExpectedImprovement.jl:59-66
    d = mp.mpf(0.5) - mp.mpf(0.01) - mp.mpf(float(mu[0])); s = mp.sqrt(mp.mpf(float(var[0]))); z = d / s
    ei_mp = d * mp.ncdf(z) + s * mp.npdf(z)
    assert abs(ei - float(ei_mp)) < 1e-9 * float(ei_mp)
    assert abs(pi - float(mp.ncdf(z))) < 1e-9 * float(mp.ncdf(z))
    assert abs(ei - 3.11345832411526e-07) < 1e-15 and ei >= 0
    assert abs(pi - 6.608138679027337e-06) < 1e-14 and 0 <= pi <= 1
    assert abs(ucb - (-1.0186125700256665)) < 1e-13
    # test/test_bayesian_opt.jl:552-558
    assert abs(ucb - (-mu[0] + 2.0 * math.sqrt(var[0]))) < 1e-10


def test_acquisition_small_variance_branches():
    # ExpectedImprovement.jl:60-62, ProbabilityImprovement.jl:58-60 (sic: max(delta, 0))
    mu = np.array([0.2, 0.9]); var = np.array([1e-12, 5e-13])
    assert np.allclose(orc.expected_improvement(mu, var, 0.0, 0.5), [0.3, 0.0])
    assert np.allclose(orc.probability_improvement(mu, var, 0.0, 0.5), [0.3, 0.0])
    assert orc.upper_confidence_bound(np.array([1.0]), np.array([-1e-3]), 2.0)[0] == -1.0


# ---- test/test_surrogates.jl:235-291  (G4)
def test_grad_kernel_blocks_se():
    x = np.array([[0.5, 0.5]]); y = np.array([[0.6, 0.6]])
    K = orc.grad_kernelmatrix(orc.SE, 1.0, 1.0, x, y)   # 3 x 3, [a, b]
    k = math.exp(-0.01)
    assert abs(K[0, 0] - k) < 1e-15
    for j in (1, 2):
        assert abs(K[0, j] - (-0.1 * k)) < 1e-14       # dk/dy_j = (x_j - y_j) k = -0.1 k
        assert abs(K[j, 0] - (0.1 * k)) < 1e-14        # dk/dx_i = -(x_i - y_i) k = +0.1 k
    for i in (1, 2):
        for j in (1, 2):
            assert abs(K[i, j] - k * ((1.0 if i == j else 0.0) - 0.01)) < 1e-14


@pytest.mark.parametrize("kind", [orc.SE, orc.APPROX_MATERN52, orc.AD_MATERN52, orc.APPROX_MATERN72, orc.AD_MATERN72])
def test_grad_kernel_blocks_vs_mp_autodiff(kind):
    """Same check as :235-291 but against an independent high-precision derivative
    (mpmath.diff) instead of ForwardDiff, for every in-repo kernel, with l = 2, sig2 = 4
    (test/test_kernels.jl:13-16)."""
    s, sc = 0.5, 4.0
    x = np.array([0.3, 0.8]); y = np.array([0.65, 0.1])

    def kmp(xa, ya):
        u = sum((mp.mpf(s) * (a - b)) ** 2 for a, b in zip(xa, ya))
        if kind == orc.SE:
            return sc * mp.e ** (-u / 2)
        r = mp.sqrt(u)
        if kind in (orc.APPROX_MATERN52, orc.AD_MATERN52):
            return sc * (1 + mp.sqrt(5) * r + 5 * u / 3) * mp.e ** (-mp.sqrt(5) * r)
        return sc * (1 + mp.sqrt(7) * r + mp.mpf(14) / 5 * u + 7 * mp.sqrt(7) / 15 * r ** 3) * mp.e ** (-mp.sqrt(7) * r)

    K = orc.grad_kernelmatrix(kind, s, sc, x[None, :], y[None, :])
    xm = [mp.mpf(float(v)) for v in x]; ym = [mp.mpf(float(v)) for v in y]
    assert abs(K[0, 0] - float(kmp(xm, ym))) < 1e-12
    for a in (1, 2):
        def fx(t, a=a):
            xx = list(xm); xx[a - 1] = t; return kmp(xx, ym)
        def fy(t, a=a):
            yy = list(ym); yy[a - 1] = t; return kmp(xm, yy)
        assert abs(K[a, 0] - float(mp.diff(fx, xm[a - 1]))) < 1e-12
        assert abs(K[0, a] - float(mp.diff(fy, ym[a - 1]))) < 1e-12
        for b in (1, 2):
            def fxy(t1, t2, a=a, b=b):
                xx = list(xm); yy = list(ym); xx[a - 1] = t1; yy[b - 1] = t2; return kmp(xx, yy)
            ref = mp.diff(fxy, (xm[a - 1], ym[b - 1]), (1, 1))
            assert abs(K[a, b] - float(ref)) < 1e-10


# ---- test/test_kernels.jl:40-63, 205-229: values == KernelFunctions Matern, incl. x == y
@pytest.mark.parametrize("kinds", [(orc.MATERN52, orc.APPROX_MATERN52, orc.AD_MATERN52),
                                   (orc.MATERN72, orc.APPROX_MATERN72, orc.AD_MATERN72)])
def test_matern_values(kinds):
    rng = np.random.default_rng(1234)
    x1, x2 = rng.random(2), rng.random(2)
    ell, sc = 2.0, 4.0
    r = np.linalg.norm(x1 - x2) / ell
    if kinds[0] == orc.MATERN52:
        ref = sc * (1 + math.sqrt(5) * r + 5 * r * r / 3) * math.exp(-math.sqrt(5) * r)
    else:
        ref = sc * (1 + math.sqrt(7) * r + 14 * r * r / 5 + 7 * math.sqrt(7) * r ** 3 / 15) * math.exp(-math.sqrt(7) * r)
    for kd in kinds:
        assert abs(orc.kernelmatrix(kd, 1 / ell, sc, x1[None], x2[None])[0, 0] - ref) < 1e-12
        assert abs(orc.kernelmatrix(kd, 1 / ell, sc, x1[None], x1[None])[0, 0] - sc) < 1e-12
    X = rng.random((5, 2))
    Ka = orc.grad_kernelmatrix(kinds[1], 1 / ell, sc, X)
    Kd = orc.grad_kernelmatrix(kinds[2], 1 / ell, sc, X)
    assert np.max(np.abs(Ka - Kd)) < 1e-12            # test_kernels.jl:57-62
    assert np.allclose(Ka, Ka.T, atol=1e-14)


# ---- test/test_kernels.jl:65-88, 231-254: gradients vs radial derivative, zero at x == y
@pytest.mark.parametrize("kind", [orc.APPROX_MATERN52, orc.AD_MATERN52, orc.APPROX_MATERN72, orc.AD_MATERN72])
def test_matern_gradients(kind):
    rng = np.random.default_rng(7)
    x1, x2 = rng.random(2), rng.random(2)
    ell, sc = 2.0, 4.0
    K = orc.grad_kernelmatrix(kind, 1 / ell, sc, x1[None], x2[None])
    r = np.linalg.norm(x1 - x2)
    h = 1e-6
    def kap(rr):
        z = rr / ell
        if kind in (orc.APPROX_MATERN52, orc.AD_MATERN52):
            return (1 + math.sqrt(5) * z + 5 * z * z / 3) * math.exp(-math.sqrt(5) * z)
        return (1 + math.sqrt(7) * z + 14 * z * z / 5 + 7 * math.sqrt(7) * z ** 3 / 15) * math.exp(-math.sqrt(7) * z)
    dk = sc * (kap(r + h) - kap(r - h)) / (2 * h)
    ref = dk * (x2 - x1) / r                            # gradient over the second argument
    assert np.max(np.abs(K[0, 1:] - ref)) < 1e-8
    K0 = orc.grad_kernelmatrix(kind, 1 / ell, sc, x1[None], x1[None])
    assert np.max(np.abs(K0[0, 1:])) < 1e-12 and np.max(np.abs(K0[1:, 0])) < 1e-12
    assert np.all(np.isfinite(K0))


# ---- test/test_surrogates.jl:293-352: GradientGP full mean and p x p covariance
def test_gradient_gp_posterior_known_answer():
    xs = np.array([[0.0, 0.0], [0.5, 0.5], [1.0, 1.0]])
    ys = np.array([[1.0, 0.1, 0.1], [0.5, 0.0, 0.0], [0.0, -0.1, -0.1]])
    noise = 0.1
    post = orc.fit_gradient(xs, ys, orc.SE, 1.0, 1.0, noise)
    xt = np.array([[0.25, 0.25]])
    gm, gv = orc.posterior_mean_var(post, xt, outputs=(0, 1, 2))
    gc = orc.posterior_cov(post, xt)

    # explicit construction, entry by entry, with independent SE block formulas
    def gk(x, a, y, b):
        k = se(x, y); D = x - y
        if a == 0 and b == 0: return k
        if a > 0 and b == 0: return -D[a - 1] * k
        if a == 0 and b > 0: return D[b - 1] * k
        return k * ((1.0 if a == b else 0.0) - D[a - 1] * D[b - 1])
    idx = [(a, i) for a in range(3) for i in range(3)]           # out-major
    Kt = np.array([[gk(xs[i], a, xs[j], b) for (b, j) in idx] for (a, i) in idx]) + noise * np.eye(9)
    yt = ys.T.reshape(-1)
    kx = np.array([[gk(xt[0], a, xs[j], b) for (b, j) in idx] for a in range(3)])  # 3 x 9
    true_mean = kx @ np.linalg.solve(Kt, yt)
    kxx = np.array([[gk(xt[0], a, xt[0], b) for b in range(3)] for a in range(3)])
    true_cov = kxx - kx @ np.linalg.solve(Kt, kx.T)
    assert np.max(np.abs(gm - true_mean)) < 1e-10
    assert np.max(np.abs(gc - true_cov)) < 1e-10
    assert np.max(np.abs(gv - np.diag(true_cov))) < 1e-10
    m1, v1 = orc.posterior_mean_var(post, xt)                      # value-only query (:985-1003)
    assert abs(m1[0] - true_mean[0]) < 1e-10 and abs(v1[0] - true_cov[0, 0]) < 1e-10 and v1[0] >= 0


# ---- test/test_kernels.jl:90-158: variance ~ 0 at a training point with zero noise
@pytest.mark.parametrize("kind", [orc.APPROX_MATERN52, orc.AD_MATERN52])
def test_gradient_gp_var_at_training_point(kind):
    rng = np.random.default_rng(1234)
    X = rng.random((5, 2))
    f = np.sin(math.pi * X[:, 0]) * np.cos(math.pi * X[:, 1])
    g = np.column_stack([math.pi * np.cos(math.pi * X[:, 0]) * np.cos(math.pi * X[:, 1]),
                         -math.pi * np.sin(math.pi * X[:, 0]) * np.sin(math.pi * X[:, 1])])
    post = orc.fit_gradient(X, np.column_stack([f, g]), kind, 0.5, 4.0, 0.0)
    m, v = orc.posterior_mean_var(post, X[:1], outputs=(0, 1, 2))
    assert abs(m[0] - f[0]) < 1e-8 and np.max(np.abs(m[1:] - g[0])) < 1e-6
    assert np.max(np.abs(v)) < 1e-6


# ---- src/acquisition_functions/acq_utils.jl:51-52
def test_sortperm_rev_semantics():
    s = np.array([1.0, 3.0, 3.0, np.nan, -1.0, 3.0, np.nan])
    assert list(orc.sortperm_rev(s)) == [3, 6, 1, 2, 5, 0, 4]
    assert list(orc.sortperm_rev(s, 3)) == [3, 6, 1]
    assert list(orc.sortperm_rev(np.array([]), 5)) == []


# ---- NLML gradient (reference: ForwardDiff through nlml, bayesian_opt.jl:284)
@pytest.mark.parametrize("kind", [orc.SE, orc.MATERN52, orc.MATERN72])
def test_nlml_gradient_standard(kind):
    rng = np.random.default_rng(3)
    X = rng.random((40, 3)); y = np.sin(3 * X).sum(1)
    th = (math.log(0.7), math.log(1.3))
    val, g = orc.nlml(X, y, kind, th[0], th[1], 1e-3, want_grad=True)
    h = 1e-5
    fd0 = (orc.nlml(X, y, kind, th[0] + h, th[1], 1e-3) - orc.nlml(X, y, kind, th[0] - h, th[1], 1e-3)) / (2 * h)
    fd1 = (orc.nlml(X, y, kind, th[0], th[1] + h, 1e-3) - orc.nlml(X, y, kind, th[0], th[1] - h, 1e-3)) / (2 * h)
    assert abs(g[0] - fd0) < 1e-5 * max(1, abs(fd0)) and abs(g[1] - fd1) < 1e-5 * max(1, abs(fd1))


def test_nlml_gradient_gradient_gp():
    rng = np.random.default_rng(4)
    X = -2 + 4 * rng.random((12, 3)); Y = orc.rosenbrock_with_grad(X) / 100.0
    yt = orc.prep_output(Y)
    th = (math.log(1.2), math.log(2.0))
    val, g = orc.nlml(X, yt, orc.APPROX_MATERN52, th[0], th[1], 1e-4, gradient_gp=True, want_grad=True)
    h = 1e-5
    f = lambda a, b: orc.nlml(X, yt, orc.APPROX_MATERN52, a, b, 1e-4, gradient_gp=True)
    fd0 = (f(th[0] + h, th[1]) - f(th[0] - h, th[1])) / (2 * h)
    fd1 = (f(th[0], th[1] + h) - f(th[0], th[1] - h)) / (2 * h)
    assert abs(g[0] - fd0) < 1e-5 * max(1, abs(fd0)) and abs(g[1] - fd1) < 1e-5 * max(1, abs(fd1))


# ---- test/test_bayesian_opt.jl:749-786: noise 0 + near-duplicate -> PosDefException
def test_posdef_failure_protocol():
    xs = np.array([[-1.0, -1.0], [5.0, -5.0], [-1.0 + 1e-12, -1.0 + 1e-12]])
    ys = np.sum(xs ** 2, axis=1)
    orc.fit_standard(xs[:2], ys[:2], orc.SE, 1.0, 1.0, 0.0)         # two points are fine
    with pytest.raises(orc.PosDefException) as ei:
        orc.fit_standard(xs, ys, orc.SE, 1.0, 1.0, 0.0)
    assert ei.value.info == 3
    with pytest.raises(ValueError):                                 # :788-817 DimensionMismatch
        orc.fit_standard(xs, ys[:2], orc.SE, 1.0, 1.0, 0.1)


def test_append_equals_refit():
    c = orc.make_config("C1", n=12, m=50)
    post = orc.fit_standard(c["X"][:11], c["y"][:11], c["kind"], c["inv_ls"], c["scale"], c["noise"])
    post2 = orc.refit_after_append(post, c["X"][11], c["y"][11])
    full = orc.fit_standard(c["X"], c["y"], c["kind"], c["inv_ls"], c["scale"], c["noise"])
    assert np.allclose(post2.alpha, full.alpha, rtol=0, atol=0)


@pytest.mark.parametrize("name,kw", [("C1", {}), ("C2", dict(n=64, m=100)), ("C3", dict(n=8, m=20, d=3)),
                                     ("C4", dict(n=64, m=100, d=5)), ("C5", dict(n=32, m=4, d=3))])
def test_configs_are_deterministic(name, kw):
    a = orc.make_config(name, **kw); b = orc.make_config(name, **kw)
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert np.array_equal(a[k], b[k])


# ---- committed golden fixture (tests/golden/known_answers.json, generated by make_known_answers.py)
def _golden():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "known_answers.json")))


def test_oracle_against_golden_fixture():
    g = _golden()
    xs = np.array([[0.0], [0.5], [1.0]])
    post = orc.fit_standard(xs, [0.0, 0.25, 1.0], orc.SE, 1.0, 1.0, 0.1)
    mu, var = orc.posterior_mean_var(post, [[0.25]])
    assert abs(mu[0] - g["G1"]["mean"]) < 1e-14 and abs(var[0] - g["G1"]["var"]) < 1e-14
    assert abs(orc.nlml(xs, [0.0, 0.25, 1.0], orc.SE, 0.0, 0.0, 0.1) - g["G2"]["nlml"]) < 1e-13
    post = orc.fit_standard(xs, [2.0, 1.0, 0.5], orc.SE, 1.0, 1.0, 0.1)
    mu, var = orc.posterior_mean_var(post, [[0.25]])
    assert abs(orc.expected_improvement(mu, var, 0.01, 0.5)[0] - g["G3"]["EI"]) < 1e-9 * g["G3"]["EI"]
    assert abs(orc.probability_improvement(mu, var, 0.01, 0.5)[0] - g["G3"]["PI"]) < 1e-9 * g["G3"]["PI"]
    assert abs(orc.upper_confidence_bound(mu, var, 2.0)[0] - g["G3"]["UCB_beta2"]) < 1e-13
    K = orc.grad_kernelmatrix(orc.SE, 1.0, 1.0, [[0.5, 0.5]], [[0.6, 0.6]])
    assert abs(K[0, 0] - g["G4"]["k"]) < 1e-15 and abs(K[0, 1] - g["G4"]["dk_dy"]) < 1e-15
    assert abs(K[1, 0] - g["G4"]["dk_dx"]) < 1e-15 and abs(K[1, 1] - g["G4"]["d2k_diag"]) < 1e-15
    assert abs(K[1, 2] - g["G4"]["d2k_offdiag"]) < 1e-15
