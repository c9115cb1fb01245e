"""Tests of the reference's own suite that had no counterpart on the CUDA path yet, reproduced through the C ABI
with the reference's inputs, assertions and tolerances:
  * standardisation equivalences           test/test_bayesian_opt.jl:238-415
  * Approx- vs AD-Matern posteriors        test/test_kernels.jl:90-158, 256-323
  * unstandardized_mean_and_var            src/surrogates/StandardGP.jl:395-404, GradientGP.jl:1019-1030
  * nlml_ls / length_scale_only            src/surrogates/StandardGP.jl:133-149, GradientGP.jl:719-738
  * DimensionMismatch out of update(BO)    test/test_bayesian_opt.jl:788-817, 858-887
The reference draws its points with Julia's RNG (Random.seed!(42) / (1234)); the assertions are equivalences and
identities that hold for any points, so NumPy draws of the same shape are used."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abo():
    import abo_b200
    return abo_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import abo_oracle
    return abo_oracle


def _f(x):
    x = np.asarray(x)
    return math.sin(x.sum()) + 0.5 * float(np.sum(x ** 2))


def _f_val_grad(x):
    x = np.asarray(x)
    return np.concatenate([[_f(x)], math.cos(x.sum()) * np.ones(len(x)) + x])


def _data():
    rng = np.random.default_rng(42)
    return [rng.standard_normal(2) for _ in range(8)]


# ---- test/test_bayesian_opt.jl:238-297 ----------------------------------------------------------------------
def test_standardgp_mean_only_equals_const_prior_mean(abo):
    x_test = _data(); y = [_f(x) for x in x_test]
    dom = abo.ContinuousDomain([-5.0, -5.0], [5.0, 5.0])
    emp = float(np.mean(y))
    k = abo.SqExponentialKernel()
    bo1 = abo.BOStruct(_f, abo.ExpectedImprovement(0.01, min(y)), abo.StandardGP(k, 1e-12), dom, x_test, y, 10, 0.0)
    bo2 = abo.BOStruct(_f, abo.ExpectedImprovement(0.01, min(y)), abo.StandardGP(k, 1e-12, mean=emp), dom, x_test, y, 10, 0.0)
    bo1, (mu1, sd1) = abo.standardize_problem(bo1, "mean_only")
    bo2.model = abo.update(bo2.model, x_test, y)
    x_pred = [[0.5, -0.3], [-1.2, 0.8], [2.1, -1.5]]
    m1 = abo.posterior_mean(bo1.model, x_pred) + mu1; v1 = abo.posterior_var(bo1.model, x_pred)
    m2 = abo.posterior_mean(bo2.model, x_pred); v2 = abo.posterior_var(bo2.model, x_pred)
    assert sd1 == 1.0 and abs(mu1 - emp) < 1e-15
    assert np.max(np.abs(m1 - m2)) < 1e-10 and np.max(np.abs(v1 - v2)) < 1e-10


# ---- test/test_bayesian_opt.jl:299-357 ----------------------------------------------------------------------
def test_standardgp_mean_scale_equals_scale_only_with_prior_mean(abo):
    x_test = _data(); y = [_f(x) for x in x_test]
    dom = abo.ContinuousDomain([-5.0, -5.0], [5.0, 5.0])
    emp = float(np.mean(y))
    k = abo.SqExponentialKernel()
    bo1 = abo.BOStruct(_f, abo.ExpectedImprovement(0.01, min(y)), abo.StandardGP(k, 1e-12), dom, x_test, y, 10, 0.0)
    bo2 = abo.BOStruct(_f, abo.ExpectedImprovement(0.01, min(y)), abo.StandardGP(k, 1e-12, mean=emp), dom, x_test, y, 10, 0.0)
    bo1, (mu1, sd1) = abo.standardize_problem(bo1, "mean_scale")
    bo2, (mu2, sd2) = abo.standardize_problem(bo2, "scale_only")
    x_pred = [[0.5, -0.3], [-1.2, 0.8], [2.1, -1.5]]
    m1 = abo.posterior_mean(bo1.model, x_pred) + mu1 / sd1; v1 = abo.posterior_var(bo1.model, x_pred)
    m2 = abo.posterior_mean(bo2.model, x_pred); v2 = abo.posterior_var(bo2.model, x_pred)
    assert mu2 == 0.0 and abs(sd1 - sd2) < 1e-15
    assert np.max(np.abs(m1 - m2)) < 1e-10 and np.max(np.abs(v1 - v2)) < 1e-10
    # standardisation also refreshes the acquisition's incumbent on the standardised ys (BO_utils.jl:61)
    assert abs(bo1.acq.best_y - min(bo1.ys)) < 1e-15


# ---- test/test_bayesian_opt.jl:359-412 ----------------------------------------------------------------------
def test_gradientgp_mean_only_equals_const_prior_mean(abo):
    x_test = _data(); Y = [_f_val_grad(x) for x in x_test]
    dom = abo.ContinuousDomain([-5.0, -5.0], [5.0, 5.0])
    emp = float(np.mean([v[0] for v in Y]))
    k = abo.SqExponentialKernel()
    best = min(v[0] for v in Y)
    bo1 = abo.BOStruct(_f_val_grad, abo.ExpectedImprovement(0.01, best), abo.GradientGP(k, 3, 1e-12), dom, x_test, Y, 10, 0.0)
    bo2 = abo.BOStruct(_f_val_grad, abo.ExpectedImprovement(0.01, best), abo.GradientGP(k, 3, 1e-12, mean=[emp, 0.0, 0.0]), dom,
                       x_test, Y, 10, 0.0)
    bo1, (mu1, sd1) = abo.standardize_problem(bo1, "mean_only")
    bo2.model = abo.update(bo2.model, x_test, Y)
    x_pred = [[0.5, -0.3], [-1.2, 0.8]]
    m1 = abo.posterior_grad_mean(bo1.model, x_pred) + np.repeat(mu1, len(x_pred))      # out-major: repeat(mu, inner = m)
    v1 = abo.posterior_grad_var(bo1.model, x_pred)
    m2 = abo.posterior_grad_mean(bo2.model, x_pred); v2 = abo.posterior_grad_var(bo2.model, x_pred)
    assert np.allclose(mu1, [emp, 0.0, 0.0], rtol=0, atol=1e-15) and np.all(sd1 == 1.0)
    assert np.max(np.abs(m1 - m2)) < 1e-10 and np.max(np.abs(v1 - v2)) < 1e-10


# ---- test/test_kernels.jl:90-158 (Matern 5/2) and :256-323 (Matern 7/2) -------------------------------------
@pytest.mark.parametrize("approx,ad", [("ApproxMatern52Kernel", "ADMatern52Kernel"), ("ApproxMatern72Kernel", "ADMatern72Kernel")])
def test_approx_vs_ad_matern_posteriors(abo, approx, ad):
    rng = np.random.default_rng(1234)
    d, ell, sig2 = 2, 2.0, 4.0
    x1 = rng.random(d)
    X = [rng.random(2) for _ in range(5)]
    f = lambda x: math.sin(math.pi * x[0]) * math.cos(math.pi * x[1])
    g = lambda x: np.array([math.pi * math.cos(math.pi * x[0]) * math.cos(math.pi * x[1]),
                            -math.pi * math.sin(math.pi * x[0]) * math.sin(math.pi * x[1])])
    Y = [np.concatenate([[f(x)], g(x)]) for x in X]
    gp_a = abo.GradientGP(sig2 * abo.with_lengthscale(abo.Kernel(approx), ell), d + 1, 0.0)
    gp_d = abo.GradientGP(sig2 * abo.with_lengthscale(abo.Kernel(ad), ell), d + 1, 0.0)
    assert abo.get_lengthscale(gp_a)[0] == abo.get_lengthscale(gp_d)[0] == 2.0       # :26-33
    assert abo.get_scale(gp_a)[0] == abo.get_scale(gp_d)[0] == 4.0
    pa = abo.update(gp_a, X, Y); pd = abo.update(gp_d, X, Y)
    for x in (x1, X[0]):                                      # a test point and a training point
        assert abs(abo.posterior_mean(pa, [x])[0] - abo.posterior_mean(pd, [x])[0]) < 1e-12
        assert np.max(np.abs(abo.posterior_grad_mean(pa, [x]) - abo.posterior_grad_mean(pd, [x]))) < 1e-10
        va, vd = abo.posterior_var(pa, [x])[0], abo.posterior_var(pd, [x])[0]
        assert abs(va - vd) < 1e-10
        assert np.max(np.abs(abo.posterior_grad_cov(pa, [x]) - abo.posterior_grad_cov(pd, [x]))) < 1e-10
    assert abo.posterior_var(pa, [x1])[0] > 0.0 and abo.posterior_var(pd, [x1])[0] > 0.0
    assert abs(abo.posterior_var(pa, [X[0]])[0]) < 1e-10 and abs(abo.posterior_var(pd, [X[0]])[0]) < 1e-10   # noise 0: interpolation
    # the posterior gradient mean at a training point reproduces the observed gradient there
    assert np.max(np.abs(abo.posterior_grad_mean(pa, [X[0]]) - Y[0])) < 1e-8


# ---- unstandardized_mean_and_var (StandardGP.jl:395-404, GradientGP.jl:1019-1030; tutorials 2D_BO.jl:177) ------
def test_unstandardized_mean_and_var(abo, orc):
    c = orc.make_config("C2", n=150, m=64)
    raw = orc.hartmann6(c["X"])
    mu, sd = float(raw.mean()), float(raw.std(ddof=1))
    k = c["scale"] * abo.with_lengthscale(abo.Matern52Kernel(), 1.0 / c["inv_ls"])
    gp = abo.update(abo.StandardGP(k, c["noise"]), c["X"], (raw - mu) / sd)
    post = orc.fit_standard(c["X"], (raw - mu) / sd, c["kind"], c["inv_ls"], c["scale"], c["noise"])
    m_o, v_o = orc.posterior_mean_var(post, c["Xc"])
    m_u, v_u = abo.unstandardized_mean_and_var(gp, c["Xc"], [mu, sd])
    assert np.max(np.abs(m_u - (m_o * sd + mu))) <= 1e-9 * max(1.0, np.max(np.abs(m_o * sd + mu)))
    assert np.max(np.abs(v_u - v_o * sd ** 2)) <= 1e-9 * sd ** 2
    # identical to rescaling the model's own posterior (the definition)
    assert np.array_equal(m_u, abo.posterior_mean(gp, c["Xc"]) * sd + mu)
    assert np.array_equal(v_u, abo.posterior_var(gp, c["Xc"]) * sd ** 2)
    # GradientGP: mean (m x p) and variance of every output, mu a p-vector (0 for the gradient outputs), one sigma
    rng = np.random.default_rng(4)
    X = -2 + 4 * rng.random((25, 3)); Y = orc.rosenbrock_with_grad(X)
    mu_g = np.array([Y[:, 0].mean(), 0.0, 0.0, 0.0]); sd_g = np.full(4, Y[:, 0].std(ddof=1))
    Ys = (Y - mu_g[None, :]) / sd_g[0]
    gg = abo.update(abo.GradientGP(1.0 * abo.with_lengthscale(abo.ApproxMatern52Kernel(), 1.5), 4, 1e-6), X, Ys)
    pg = orc.fit_gradient(X, Ys, 3, 1.0 / 1.5, 1.0, 1e-6)
    Xq = -2 + 4 * rng.random((9, 3))
    mo, vo = orc.posterior_mean_var(pg, Xq, outputs=range(4))
    mu_u, v_uu = abo.unstandardized_mean_and_var(gg, Xq, (mu_g, sd_g))
    ref_m = mo.reshape(4, -1).T * sd_g[0] + mu_g[None, :]; ref_v = vo.reshape(4, -1).T * sd_g[0] ** 2
    assert mu_u.shape == (9, 4) and np.max(np.abs(mu_u - ref_m)) <= 1e-9 * np.max(np.abs(ref_m))
    assert np.max(np.abs(v_uu - ref_v)) <= 1e-9 * max(1.0, np.max(np.abs(ref_v)))


# ---- nlml_ls and hyper-parameter optimisation with length_scale_only (StandardGP.jl:133-149, bayesian_opt.jl:259) ----
def test_nlml_ls_and_length_scale_only(abo, orc):
    rng = np.random.default_rng(8)
    X = rng.random((80, 3)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
    gp = abo.StandardGP(2.0 * abo.with_lengthscale(abo.Matern52Kernel(), 0.4), 1e-4, mean=0.2)
    for log_ls, log_sc in ((math.log(0.4), math.log(2.0)), (0.3, -0.7)):
        v = abo.nlml_ls(gp, log_ls, log_sc, X, y)
        assert v == abo.nlml(gp, [log_ls, log_sc], X, y)
        v_o = orc.nlml(X, y, 1, log_ls, log_sc, 1e-4, mean_c=0.2)
        assert abs(v - v_o) <= 1e-9 * abs(v_o)
    old = [math.log(0.4), math.log(2.0)]
    new = abo.optimize_hyperparameters(gp, X, y, old, length_scale_only=True, num_restarts=3, rng=rng)
    assert isinstance(new, abo.StandardGP) and abo.get_scale(new)[0] == 2.0          # the scale is frozen
    assert abo.nlml_ls(new, math.log(abo.get_lengthscale(new)[0]), math.log(2.0), X, y) <= abo.nlml_ls(gp, old[0], old[1], X, y) + 1e-6
    # GradientGP flavour (GradientGP.jl:719-738)
    Xg = -2 + 4 * rng.random((12, 2)); Yg = orc.rosenbrock_with_grad(Xg) / 50.0
    gg = abo.GradientGP(1.0 * abo.with_lengthscale(abo.ApproxMatern52Kernel(), 1.0), 3, 1e-4)
    vg = abo.nlml_ls(gg, 0.1, -0.2, Xg, Yg)
    vo = orc.nlml(Xg, orc.prep_output(Yg), 3, 0.1, -0.2, 1e-4, gradient_gp=True)
    assert abs(vg - vo) <= 1e-9 * abs(vo)


# ---- test/test_bayesian_opt.jl:858-887: GradientGP flavour of the wrong-dimension update ------------------------
def test_update_bo_wrong_dimension_gradient_gp(abo):
    f = lambda x: np.concatenate([[float(np.sum(np.asarray(x) ** 2))], 2 * np.asarray(x)])
    dom = abo.ContinuousDomain([-2.0, -2.0], [2.0, 2.0])
    xs = [[-1.0, -1.0], [1.5, -0.5]]; ys = [f(x) for x in xs]
    gp = abo.update(abo.GradientGP(abo.SqExponentialKernel(), 3, 0.1), xs, ys)
    bo = abo.BOStruct(f, abo.ExpectedImprovement(0.01, 2.0), gp, dom, xs, ys, 10, 0.0)
    with pytest.raises(abo.DimensionMismatch):
        abo.update(bo, [0.0], f([0.0]), 0)


# ---- get_mean_std / std_y on the device (BO_utils.jl:44-64, StandardGP.jl:164-204, GradientGP.jl:756-783) ----------
@pytest.mark.parametrize("choice", ["mean_scale", "scale_only", "mean_only"])
def test_device_standardisation_matches_reference_formulas(abo, choice):
    rng = np.random.default_rng(11)
    ctx = abo.default_context()
    y = 3.0 + 2.5 * rng.standard_normal(5000)
    mu, sd, ys, best = ctx.standardize(y, len(y), 1, choice)
    mu_ref = 0.0 if choice == "scale_only" else float(np.mean(y))
    sd_ref = 1.0 if choice == "mean_only" else float(np.std(y, ddof=1))          # Statistics.std: corrected
    assert abs(mu[0] - mu_ref) <= 1e-14 * max(1.0, abs(mu_ref)) and abs(sd[0] - sd_ref) <= 1e-14 * sd_ref
    assert np.max(np.abs(ys - (y - mu_ref) / sd_ref)) <= 1e-13 and best == ys.min()
    # GradientGP: only the value output is centred, every output is divided by the value output's std
    Y = np.column_stack([y[:400], rng.standard_normal((400, 3))])
    mu_g, sd_g, ys_g, best_g = ctx.standardize(Y.T.reshape(-1), 400, 4, choice)
    m0 = 0.0 if choice == "scale_only" else float(np.mean(Y[:, 0])); s0 = 1.0 if choice == "mean_only" else float(np.std(Y[:, 0], ddof=1))
    assert np.allclose(mu_g, [m0, 0, 0, 0], rtol=0, atol=1e-14) and np.allclose(sd_g, s0, rtol=1e-14, atol=0)
    ref = (Y - np.array([m0, 0, 0, 0])[None, :]) / s0
    assert np.max(np.abs(ys_g.reshape(4, 400).T - ref)) <= 1e-13 and best_g == ys_g[:400].min()
    # the host helpers the reference's loop calls give the same numbers
    g = abo.StandardGP(abo.SqExponentialKernel(), 1e-6)
    mu_h, sd_h = abo.get_mean_std(g, y, choice)
    assert abs(mu_h - mu[0]) <= 1e-14 * max(1.0, abs(mu_h)) and abs(sd_h - sd[0]) <= 1e-14 * sd_h


# ---- the reference's own standardisation vectors (test/test_surrogates.jl:107-128, 366-397) on the device kernel and on
#      the host helpers of the mirror -----------------------------------------------------------------------------------
def test_reference_standardisation_known_answers(abo):
    ctx = abo.default_context()
    y_train = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    mu, sd, ys, best = ctx.standardize(y_train, 5, 1, "mean_scale")
    assert abs(mu[0] - 3.0) <= 1e-15 and sd[0] > 0                       # "μ ≈ 3.0", "σ > 0"
    assert abs(np.std(ys, ddof=1) - 1.0) <= 1e-10                        # "std(y_flat_std) ≈ 1.0 atol = 1e-10"
    g = abo.StandardGP(abo.SqExponentialKernel(), 0.1)
    mu_h, sd_h = abo.get_mean_std(g, y_train, "mean_scale")
    assert mu_h == mu[0] and abs(sd_h - sd[0]) <= 1e-15 and np.allclose(abo.std_y(g, y_train, mu_h, sd_h), ys, rtol=0, atol=1e-15)
    # GradientGP: y_train = [[1, .1, .1], [2, .2, .2], [3, .3, .3]]
    Y = np.array([[1.0, 0.1, 0.1], [2.0, 0.2, 0.2], [3.0, 0.3, 0.3]])
    mu_g, sd_g, ys_g, _ = ctx.standardize(Y.T.reshape(-1), 3, 3, "mean_scale")         # out-major, like prep_output
    assert len(mu_g) == 3 and len(sd_g) == 3
    assert abs(mu_g[0] - 2.0) <= 1e-15 and mu_g[1] == 0.0 and mu_g[2] == 0.0            # gradients keep a zero mean
    assert sd_g[0] > 0 and sd_g[1] == sd_g[0] and sd_g[2] == sd_g[0]                    # and the value output's scale
    Ys = ys_g.reshape(3, 3).T
    for y_o, y_s in zip(Y, Ys):
        for a in range(3):
            assert abs(y_s[a] - (y_o[a] - mu_g[a]) / sd_g[a]) <= 1e-8
    gg = abo.GradientGP(abo.SqExponentialKernel(), 3, 0.1)
    mu_hh, sd_hh = abo.get_mean_std(gg, Y, "mean_scale")
    assert np.allclose(mu_hh, mu_g, rtol=0, atol=1e-15) and np.allclose(sd_hh, sd_g, rtol=1e-15, atol=0)
    assert np.allclose(abo.std_y(gg, Y, mu_hh, sd_hh), Ys, rtol=0, atol=1e-15)
