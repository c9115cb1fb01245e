"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/abo.h
declares, and fails loudly (no fallback) without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import abo_b200
    return abo_b200


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "abo.h")).read()
    declared = set(re.findall(r"(?:int32_t|const char\*)\s+(abo_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(built.SYMBOLS), declared ^ set(built.SYMBOLS)
    L = ctypes.CDLL(built.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    L.abo_version.restype = ctypes.c_int32
    assert L.abo_version() >= 100


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built.AboCudaError) as e:
        built.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "abstractbayesopt.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, f"{f} references the oracle"


def test_host_kernel_algebra(built):
    abo = built
    k = 4.0 * abo.with_lengthscale(abo.Matern52Kernel(), 2.0)
    gp = abo.StandardGP(k, 0.1)
    assert abo.get_lengthscale(gp) == [2.0] and abo.get_scale(gp) == [4.0]       # test_kernels.jl:27-34
    gp2 = abo.StandardGP(abo.SqExponentialKernel(), 0.1)
    assert abo.get_lengthscale(gp2) == [1.0] and abo.get_scale(gp2) == [1.0]     # test_surrogates.jl:30-57
    gp3 = abo.StandardGP(3.0 * abo.SqExponentialKernel(), 0.1)
    assert abo.get_lengthscale(gp3) == [1.0] and abo.get_scale(gp3) == [3.0]
    assert gp.gpx is None
    r = abo.rescale_model(gp, 2.0)
    assert abo.get_scale(r) == [1.0] and abs(r.noise_var - 0.025) < 1e-15
    with pytest.raises(ValueError):
        abo.ContinuousDomain([0.0, 1.0], [1.0, 0.0])                              # test_domains.jl
    d = abo.ContinuousDomain([0.0, 0.0], [1.0, 2.0])
    g = abo.latin_hypercube(100, d.lower, d.upper, __import__("numpy").random.default_rng(0))
    assert g.shape == (100, 2) and (g >= d.lower).all() and (g <= d.upper).all()
    import numpy as np
    for k in range(2):   # one point per stratum
        assert sorted(np.floor((g[:, k] - d.lower[k]) / (d.upper[k] - d.lower[k]) * 100).astype(int)) == list(range(100))


def test_standardisation_helpers(built):
    import numpy as np
    abo = built
    gp = abo.StandardGP(abo.SqExponentialKernel(), 0.1)
    ys = np.array([1.0, 2.0, 4.0])
    mu, sd = abo.get_mean_std(gp, ys, "mean_scale")
    assert abs(mu - ys.mean()) < 1e-15 and abs(sd - ys.std(ddof=1)) < 1e-15
    assert abo.get_mean_std(gp, ys, "scale_only")[0] == 0.0 and abo.get_mean_std(gp, ys, "mean_only")[1] == 1.0
    ggp = abo.GradientGP(abo.ApproxMatern52Kernel(), 3, 0.1)
    Y = np.array([[1.0, 0.1, 0.2], [3.0, 0.3, 0.1], [2.0, -0.2, 0.0]])
    mu, sd = abo.get_mean_std(ggp, Y, "mean_scale")
    assert mu[0] == 2.0 and mu[1] == 0 and mu[2] == 0 and sd[0] == sd[1] == sd[2] == 1.0
    assert np.array_equal(abo.prep_output(ggp, Y), Y.T.reshape(-1))


def test_lockstep_lbfgsb_on_analytic_problems(built):
    """The lock-step multi-start optimiser (host logic of optimize_hyperparameters) on problems with
    known minimisers, including active box constraints and a restart whose start is infeasible."""
    import numpy as np
    abo = built
    rng = np.random.default_rng(0)
    R = 9
    centers = rng.uniform(-1, 1, (R, 2)); centers[3] = [3.0, 0.0]       # minimiser outside the box -> on the bound
    calls = []

    def fg(X, idx):
        calls.append(len(idx))
        d = X[idx] - centers[idx]
        val = 0.5 * (d[:, 0] ** 2 + 10 * d[:, 1] ** 2) + 0.1 * d[:, 0] ** 4
        grad = np.column_stack([d[:, 0] + 0.4 * d[:, 0] ** 3, 10 * d[:, 1]])
        val = np.where(idx == 5, np.inf, val)                              # restart 5 always fails
        return val, grad

    x0 = rng.uniform(-2, 2, (R, 2))
    X, f, conv, failed = abo.lockstep_lbfgsb(fg, x0, np.array([-2.0, -2.0]), np.array([2.0, 2.0]))
    assert failed[5] and not failed[[0, 1, 2, 3, 4, 6, 7, 8]].any()
    ok = [0, 1, 2, 4, 6, 7, 8]
    assert conv[ok].all() and np.max(np.abs(X[ok] - centers[ok])) < 1e-4
    assert abs(X[3, 0] - 2.0) < 1e-12 and abs(X[3, 1]) < 1e-4              # clipped to the upper bound
    assert max(calls) <= R and len(calls) < 200                            # batched: one call per trial step


def test_lockstep_lbfgsb_rosenbrock_with_history(built):
    """Curved valleys need the quasi-Newton history (vectorised two-loop recursion): 40 Rosenbrock problems from
    random starts, free and with an active upper bound."""
    import numpy as np
    abo = built

    def fg(X, idx):
        x = X[idx]
        f = (1 - x[:, 0]) ** 2 + 100 * (x[:, 1] - x[:, 0] ** 2) ** 2
        g = np.stack([-2 * (1 - x[:, 0]) - 400 * x[:, 0] * (x[:, 1] - x[:, 0] ** 2), 200 * (x[:, 1] - x[:, 0] ** 2)], 1)
        return f, g

    x0 = np.random.default_rng(0).uniform(-2, 2, (40, 2))
    X, f, conv, failed = abo.lockstep_lbfgsb(fg, x0, np.array([-2.0, -2.0]), np.array([2.0, 2.0]), max_iter=300, g_tol=1e-8,
                                             f_abstol=0.0)
    assert conv.all() and not failed.any() and np.max(np.abs(X - 1.0)) < 1e-6
    X, f, conv, failed = abo.lockstep_lbfgsb(fg, x0, np.array([-2.0, -2.0]), np.array([0.5, 2.0]), max_iter=300, g_tol=1e-8,
                                             f_abstol=0.0)
    assert conv.all() and np.max(np.abs(X - np.array([0.5, 0.25]))) < 1e-6 and np.max(np.abs(f - 0.25)) < 1e-10
