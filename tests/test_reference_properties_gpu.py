"""The reference's "Mathematical Properties and Correctness" and "Numerical Stability" testsets
(test/test_bayesian_opt.jl:458-745, 889-940) with the same inputs and thresholds, through the CUDA path."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abo():
    import abo_b200
    return abo_b200


def test_gp_posterior_consistency(abo):                                     # :461-487
    f = lambda x: x[0] ** 2 + 0.5 * x[1] ** 2
    x_train = [[-1.0, -1.0], [0.0, 0.0], [1.0, 1.0]]
    y_train = [f(x) for x in x_train]
    gp = abo.update(abo.StandardGP(abo.SqExponentialKernel(), 0.01), x_train, y_train)
    assert np.all(np.abs(abo.posterior_mean(gp, x_train) - np.array(y_train)) < 0.1)
    assert np.all(abo.posterior_var(gp, x_train) < 0.1)
    v = abo.posterior_var(gp, [[0.1, 0.1], [1.5, 1.5], [3.0, 3.0]])       # close, medium, far
    assert v[0] < v[1] < v[2]


def test_acquisition_mathematical_properties(abo):                          # :512-559
    kernel = 1.0 * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0)
    x_train, y_train = [-1.0, 0.0, 1.0], [1.0, 0.25, 1.0]
    gp = abo.update(abo.StandardGP(kernel, 0.01), x_train, y_train)
    ei = abo.ExpectedImprovement(0.01, min(y_train))
    assert np.all(ei(gp, [-2.0, -0.5, 0.5, 2.0]) >= 0.0)
    noiseless = abo.update(abo.StandardGP(kernel, 1e-12), x_train, y_train)
    assert ei(noiseless, [x_train[int(np.argmin(y_train))]])[0] < 0.01
    test_x = [0.5]
    assert abo.UpperConfidenceBound(1.0)(gp, test_x)[0] < abo.UpperConfidenceBound(3.0)(gp, test_x)[0]
    ucb_val = abo.UpperConfidenceBound(2.0)(gp, test_x)[0]
    expected = -abo.posterior_mean(gp, test_x)[0] + 2.0 * math.sqrt(abo.posterior_var(gp, test_x)[0])
    assert abs(ucb_val - expected) < 1e-10


def test_optimisation_improves_monotonically(abo):                          # :561-595
    f = lambda x: (float(np.ravel(x)[0]) - 0.7) ** 2 + 0.1
    domain = abo.ContinuousDomain([-2.0], [3.0])
    x_train = [-1.0, 0.0, 2.0]
    y_train = [f(x) for x in x_train]
    gp = abo.StandardGP(1.0 * abo.with_lengthscale(abo.SqExponentialKernel(), 2.0), 0.01)       # ScaleTransform(0.5)
    problem = abo.BOStruct(f, abo.ExpectedImprovement(0.01, min(y_train)), gp, domain, x_train, y_train, 15, 0.01)
    result, _, _ = abo.optimize(problem, standardize=None, hyper_params=None, rng=np.random.default_rng(42))
    ys = np.array([float(np.ravel(v)[0]) for v in result.ys_non_std])
    best = np.minimum.accumulate(ys)
    assert np.all(np.diff(best) <= 0) and len(ys) > 3
    assert best[-1] < min(y_train)                                           # and it does find something better than the start


def test_hyperparameter_optimisation_consistency(abo):                      # :597-653
    rng = np.random.default_rng(42)
    x_train = [-1.0, -0.5, 0.0, 0.5, 1.0]
    y_train = [math.sin(2 * x) + 0.1 * rng.standard_normal() for x in x_train]
    gp = abo.update(abo.StandardGP(abo.SqExponentialKernel(), 0.01), x_train, y_train)
    initial = [math.log(1.0), math.log(1.0)]
    opt = abo.optimize_hyperparameters(gp, x_train, y_train, initial, num_restarts=3, rng=rng)
    ls, sc = abo.get_lengthscale(opt)[0], abo.get_scale(opt)[0]
    assert ls > 0 and sc > 0
    assert abo.nlml(opt, [math.log(ls), math.log(sc)], x_train, y_train) <= abo.nlml(gp, initial, x_train, y_train) + 1e-6


def test_gradient_gp_mathematical_consistency(abo):                         # :655-705
    fvg = lambda x: [x[0] ** 2 + x[1] ** 2, 2 * x[0], 2 * x[1]]
    x_train = [[-1.0, -1.0], [0.0, 0.0], [1.0, 1.0]]
    y_train = [fvg(x) for x in x_train]
    gp = abo.update(abo.GradientGP(abo.SqExponentialKernel(), 3, 1e-12), x_train, y_train)
    test_x = [[0.5, -0.3]]
    pred_f = abo.posterior_mean(gp, test_x)[0]
    pred_full = abo.posterior_grad_mean(gp, test_x)
    assert abs(pred_f - pred_full[0]) < 1e-10 and len(pred_full) == 3
    for i, x in enumerate(x_train):
        at = abo.posterior_grad_mean(gp, [x])
        for j in range(3):
            assert abs(at[j] - y_train[i][j]) < 1e-4


def test_standardisation_mathematical_correctness(abo):                     # :707-745
    f = lambda x: 3 * float(np.ravel(x)[0]) ** 2 + 5.0
    x_train = [-1.0, -0.5, 0.0, 0.5, 1.0]
    y_train = [f(x) for x in x_train]
    emp_mean, emp_std = float(np.mean(y_train)), float(np.std(y_train, ddof=1))
    gp = abo.StandardGP(abo.SqExponentialKernel(), 0.01)
    domain = abo.ContinuousDomain([-2.0], [2.0])
    mk = lambda: abo.BOStruct(f, abo.ExpectedImprovement(0.01, min(y_train)), gp, domain, x_train, y_train, 5, 0.01)
    std_problem, params = abo.standardize_problem(mk(), "mean_scale")
    mu, sd = params
    assert abs(float(np.ravel(sd)[0]) - emp_std) < 1e-10 and abs(float(np.ravel(mu)[0]) - emp_mean) < 1e-10
    recovered = abo.rescale_output(std_problem.ys, params)
    assert np.all(np.abs(np.array(recovered, dtype=np.float64).reshape(-1) - np.array(y_train)) < 1e-10)
    std_scale, _ = abo.standardize_problem(mk(), "scale_only")
    assert abs(float(np.mean(np.array(std_scale.ys, dtype=np.float64))) - emp_mean / emp_std) < 0.1


def test_near_singular_kernel_matrix(abo):                                  # :889-911
    gp = abo.StandardGP(abo.SqExponentialKernel(), 1e-12)
    try:
        up = abo.update(gp, [0.0, 1e-10, 2e-10], [1.0, 1.001, 1.002])
        assert math.isfinite(abo.posterior_mean(up, [0.5])[0])
    except abo.PosDefException:                                             # the reference accepts PosDef / Singular here
        pass


def test_extreme_hyperparameter_values(abo):                                # :913-940
    x_train, y_train = [-1.0, 0.0, 1.0], [1.0, 0.0, 1.0]
    large = abo.update(abo.StandardGP(abo.SqExponentialKernel(), 0.01), x_train, y_train)
    pred_large = abo.posterior_mean(large, [0.5])[0]
    small = abo.update(abo.StandardGP(1.0 * abo.with_lengthscale(abo.SqExponentialKernel(), 1e-6), 0.01), x_train, y_train)   # ScaleTransform(1e6)
    pred_small = abo.posterior_mean(small, [0.5])[0]
    assert math.isfinite(pred_large) and math.isfinite(pred_small) and abs(pred_large - pred_small) > 0.01


# ---- "BOStruct Tests" (test/test_bayesian_opt.jl:9-223) ---------------------------------------------------------------
def test_bostruct_construction_update_and_utilities(abo, capsys):
    f = lambda x: float(np.sum(np.asarray(x) ** 2))
    domain = abo.ContinuousDomain([-2.0, -2.0], [2.0, 2.0])
    gp = abo.StandardGP(abo.SqExponentialKernel(), 0.1)
    x_train = [[-1.0, -1.0], [0.0, 0.0], [1.0, 1.0]]
    y_train = [f(x) for x in x_train]
    problem = abo.BOStruct(f, abo.ExpectedImprovement(0.01, min(y_train)), gp, domain, x_train, y_train, 10, 0.1)
    assert problem.func is f and problem.domain is domain and problem.max_iter == 10 and problem.noise == 0.1      # :10-41
    assert problem.iter == 0 and problem.flag is False and len(problem.xs) == 3 and len(problem.ys) == 3
    x_train = [[-1.0, -1.0], [5.0, -5.0]]                                                                          # :43-79
    y_train = [f(x) for x in x_train]
    up_gp = abo.update(gp, x_train, y_train)
    problem = abo.BOStruct(f, abo.ExpectedImprovement(0.01, min(y_train)), up_gp, domain, x_train, y_train, 10, 0.1)
    x_new = [0.0, 0.0]
    up = abo.update(problem, x_new, f(x_new), 0)
    assert len(up.xs) == 3 and len(up.ys) == 3 and list(up.xs[-1]) == x_new and float(up.ys[-1]) == f(x_new)
    assert up.iter == 1 and up.acq.best_y == 0.0
    abo.print_info(up)                                                                                             # :81-106
    out = capsys.readouterr().out
    assert "== BOStruct Information ==" in out and "Number of data points: 3" in out and "Max iterations: 10" in out


def test_hyperparameter_optimisation_returns_a_surrogate(abo):              # :108-137
    gp = abo.update(abo.StandardGP(abo.SqExponentialKernel(), 0.1), [[-1.0], [0.0], [1.0]], [1.0, 0.0, 1.0])
    opt = abo.optimize_hyperparameters(gp, [[-1.0], [0.0], [1.0]], [1.0, 0.0, 1.0], [0.0, 0.0], num_restarts=2, scale_std=1.0,
                                       rng=np.random.default_rng(0))
    assert isinstance(opt, abo.StandardGP)


@pytest.mark.parametrize("mode", ["mean_scale", "scale_only", "mean_only"])
def test_standardisation_modes(abo, mode):                                  # :139-184
    f = lambda x: float(np.sum(np.asarray(x) ** 2)) + 10.0
    domain = abo.ContinuousDomain([-2.0], [2.0])
    gp = abo.StandardGP(abo.SqExponentialKernel(), 0.1)
    x_train = [[-1.0], [0.0], [1.0]]
    y_train = [f(x) for x in x_train]
    problem = abo.BOStruct(f, abo.ExpectedImprovement(0.01, min(y_train)), gp, domain, x_train, y_train, 10, 0.1)
    std_problem, params = abo.standardize_problem(problem, mode)
    mu, sd = params
    assert float(np.ravel(sd)[0]) > 0 and isinstance(std_problem.model, abo.StandardGP)
    if mode in ("scale_only", "mean_scale"):
        assert len(abo.rescale_output(std_problem.ys, params)) == len(problem.ys_non_std)


def test_optimisation_loop(abo):                                            # :186-223
    f = lambda x: (float(np.ravel(x)[0]) - 0.5) ** 2
    domain = abo.ContinuousDomain([-1.0], [2.0])
    x_train = [-0.5, 0.0, 1.5]
    y_train = [f(x) for x in x_train]
    gp = abo.StandardGP(abo.SqExponentialKernel(), 0.01)
    problem = abo.BOStruct(f, abo.ExpectedImprovement(0.01, min(y_train)), gp, domain, x_train, y_train, 3, 0.01)
    result, acq_list, std_params = abo.optimize(problem, standardize=None, hyper_params=None, rng=np.random.default_rng(3))
    assert len(result.xs) >= len(x_train) and len(result.ys) >= len(y_train) and len(acq_list) >= 0 and result.iter > 0
    assert min(float(np.ravel(v)[0]) for v in result.ys_non_std) <= min(y_train)


# ---- the "Evaluation" testsets of test/test_acquisition.jl (:20-43, :74-95, :126-149, :159-181, :223-253) ----------------
def test_acquisition_evaluation_vectors(abo):
    gp = abo.update(abo.StandardGP(abo.SqExponentialKernel(), 0.1), [0.0, 0.5, 1.0], [2.0, 1.0, 0.5])
    ei, pi, ucb = abo.ExpectedImprovement(0.01, 0.5), abo.ProbabilityImprovement(0.01, 0.5), abo.UpperConfidenceBound(2.0)
    ei_val = ei(gp, [0.25])[0]
    assert math.isfinite(ei_val) and ei_val >= 0.0
    ucb_val = ucb(gp, 0.25)[0]                                               # a bare Real is one 1-D point
    assert math.isfinite(ucb_val)
    pi_val = pi(gp, 0.25)[0]
    assert math.isfinite(pi_val) and 0.0 <= pi_val <= 1.0
    ens = abo.EnsembleAcquisition([0.6, 0.4], [ei, ucb])
    ens_val = ens(gp, 0.25)
    assert np.all(np.isfinite(ens_val))
    assert np.allclose(ens_val, 0.6 * ei(gp, 0.25) + 0.4 * ucb(gp, 0.25), rtol=1e-12, atol=0)
    ggp = abo.update(abo.GradientGP(abo.SqExponentialKernel(), 3, 0.1), [[0.0, 0.0], [0.5, 0.5], [1.0, 1.0]],
                     [[2.0, 0.1, 0.1], [1.0, 0.0, 0.0], [0.5, -0.1, -0.1]])
    g_val = abo.GradientNormUCB(2.0)(ggp, [[0.25, 0.25]])[0]
    assert math.isfinite(g_val)
