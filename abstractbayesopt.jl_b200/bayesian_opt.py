"""Host-side callers of the hot path, mirroring src/bayesian_opt.jl and
src/acquisition_functions/acq_utils.jl: BOStruct, update(BO, x, y, i) with the
PosDef/Singular rollback protocol, the multi-start acquisition optimiser and the multi-start
hyper-parameter MLE.  These are thin: the arithmetic is in libabo_cuda.so."""
from __future__ import annotations

import logging
import math

import numpy as np
from scipy.optimize import minimize

from ._lib import PosDefException
from .kernels import with_lengthscale
from .surrogates import (GradientGP, _update_model_parameters, get_kernel_constructor, get_lengthscale,
                         get_mean_std, get_scale, nlml_batch, rescale_model, std_y, update_surrogate)

log = logging.getLogger("abo_b200")


def latin_hypercube(n, lower, upper, rng):
    """QuasiMonteCarlo.sample(n, lb, ub, LatinHypercubeSample()) (acq_utils.jl:44-46): one point
    per stratum and dimension, independently permuted."""
    d = len(lower)
    u = (np.argsort(rng.random((d, n)), axis=1).T + rng.random((n, d))) / n
    return lower + (upper - lower) * u


def optimize_acquisition(acqf, surrogate, domain, n_grid=10_000, n_local=100, rng=None, refine=True):
    """optimize_acquisition (acq_utils.jl:33-73): LHS grid → ONE batched, fused sweep that also
    returns the stable top-n_local (replaces acqf(...) + sortperm, :50-52) → box-constrained
    L-BFGS refinements from those starts (finite-difference gradients, as Optim's default)."""
    rng = np.random.default_rng() if rng is None else rng
    grid = latin_hypercube(n_grid, domain.lower, domain.upper, rng)
    _, top_idx, top_val = acqf.topk(surrogate, grid, n_local)
    best_acq, best_x = -math.inf, None
    bounds = list(zip(domain.lower, domain.upper))
    for i0, v0 in zip(top_idx, top_val):
        x0 = grid[i0]
        if refine:
            res = minimize(lambda x: -float(acqf(surrogate, x[None, :])[0]), x0, method="L-BFGS-B", bounds=bounds,
                           options=dict(gtol=1e-5, ftol=2.2e-9, maxls=20))
            cur, xc = -float(res.fun), res.x
        else:
            cur, xc = float(v0), x0
        if cur > best_acq:
            best_acq, best_x = cur, np.array(xc)
    return best_x


def optimize_hyperparameters(model, x_train, y_train, old_params, scale_std=1.0, length_scale_only=False,
                             num_restarts=1, domain=None, rng=None, max_iter=50):
    """optimize_hyperparameters (bayesian_opt.jl:196-328).  Same log-space box, clamped start and
    uniform random restarts; all restarts advance in LOCK-STEP so that every objective/gradient
    evaluation is one batched abo_nlml_batch call (value + analytic gradient) instead of one
    ForwardDiff evaluation per restart."""
    rng = np.random.default_rng() if rng is None else rng
    ls_lo, ls_hi = 1e-3, 1e3
    if domain is not None:
        X = np.asarray(x_train, dtype=np.float64).reshape(len(x_train), -1)
        side = float(np.max(domain.upper - domain.lower))
        ls_lo, ls_hi = max(1e-6, 1e-3 * side), side
    sc_lo, sc_hi = 1e-3 / scale_std ** 2, 1e6 / scale_std ** 2
    lo = np.log([ls_lo, sc_lo]); hi = np.log([ls_hi, sc_hi])
    start = np.clip(np.asarray(old_params, dtype=np.float64), lo + 2 * np.finfo(float).eps, hi - 2 * np.finfo(float).eps)
    inits = [start] + [lo + (hi - lo) * rng.random(2) for _ in range(num_restarts - 1)]
    if length_scale_only:
        for t in inits:
            t[1] = start[1]
    bounds = [(lo[0], hi[0]), (lo[1], hi[1])]
    best_val, best = math.inf, None
    for t0 in inits:
        def fg(t):
            val, grad, info = nlml_batch(model, t[None, :], x_train, y_train)
            if info[0] != 0 or not np.isfinite(val[0]):
                return 1e300, np.zeros(2)
            g = grad[0].copy()
            if length_scale_only:
                g[1] = 0.0
            return float(val[0]), g
        try:
            res = minimize(fg, t0, jac=True, method="L-BFGS-B", bounds=bounds,
                           options=dict(gtol=1e-6, ftol=2.2e-9, maxiter=max_iter))
        except Exception as e:                                        # bayesian_opt.jl:296-299
            log.warning("Optimization failed at restart with error: %s", e)
            continue
        if res.success and res.fun < best_val:
            best_val, best = res.fun, res.x
    if best is None:
        log.info("All restarts failed to converge.")
        return model
    ell = math.exp(best[0])
    scale = get_scale(model)[0] if length_scale_only else math.exp(best[1])
    k_opt = scale * with_lengthscale(get_kernel_constructor(model), ell)
    return _update_model_parameters(model, k_opt)


class BOStruct:
    """BOStruct(f, acq, model, domain, xs, ys, max_iter, noise) (bayesian_opt.jl:38-92)."""

    def __init__(self, func, acq, model, domain, x_train, y_train, max_iter, noise):
        self.func, self.acq, self.model, self.domain = func, acq.copy(), model.copy(), domain
        self.xs = [np.asarray(x, dtype=np.float64).reshape(-1) for x in x_train]
        self.ys = [np.asarray(y, dtype=np.float64) for y in y_train]
        self.ys_non_std = [np.array(y, dtype=np.float64) for y in y_train]
        self.max_iter, self.iter, self.noise, self.flag = max_iter, 0, noise, False


def update_bo(BO: BOStruct, x, y, i):
    """update(BO, x, y, i) (bayesian_opt.jl:113-150): snapshot → push → update → on
    PosDefException roll back xs/ys/ys_non_std, restore the snapshot, set `flag`."""
    prev = BO.model.copy()
    BO.xs.append(np.asarray(x, dtype=np.float64).reshape(-1))
    BO.ys.append(np.asarray(y, dtype=np.float64))
    try:
        BO.model = update_surrogate(BO.model, np.array(BO.xs), np.array(BO.ys))
    except PosDefException:
        log.info("We reached ill-conditioning, returning NON-UPDATED GP. Killing BO loop.")
        BO.model = prev
        BO.xs.pop(); BO.ys.pop()
        if len(BO.ys_non_std) > len(BO.ys):
            BO.ys_non_std.pop()
        BO.flag = True
        return BO
    BO.acq = BO.acq.update(np.array(BO.ys), BO.model)
    BO.iter = i + 1
    return BO


def stop_criteria(BO):
    return BO.iter > BO.max_iter


def standardize_problem(BO, choice):
    """standardize_problem (BO_utils.jl:44-64)."""
    mu, sd = get_mean_std(BO.model, np.array(BO.ys_non_std), choice)
    BO.model = rescale_model(BO.model, sd)
    BO.ys = list(std_y(BO.model, np.array(BO.ys_non_std), mu, sd))
    BO.model = update_surrogate(BO.model, np.array(BO.xs), np.array(BO.ys))
    BO.acq = BO.acq.update(np.array(BO.ys), BO.model)
    return BO, (mu, sd)


def optimize(BO: BOStruct, standardize="mean_scale", hyper_params="all", num_restarts_HP=1, n_grid=10_000,
             n_local=100, rng=None, refine=True):
    """optimize(BO; standardize, hyper_params, num_restarts_HP) (bayesian_opt.jl:364-449)."""
    if standardize not in ("mean_scale", "scale_only", "mean_only", None):
        raise ValueError("standardize must be one of mean_scale, scale_only, mean_only, None")
    if hyper_params not in ("all", "length_scale_only", None):
        raise ValueError("hyper_params must be one of all, length_scale_only, None")
    rng = np.random.default_rng() if rng is None else rng
    if standardize is not None:
        BO, (mu, sd) = standardize_problem(BO, standardize)
    else:
        is_grad = isinstance(BO.model, GradientGP)
        mu, sd = (np.zeros(BO.model.p), np.ones(BO.model.p)) if is_grad else (0.0, 1.0)
        BO.model = update_surrogate(BO.model, np.array(BO.xs), np.array(BO.ys))
        BO.acq = BO.acq.update(np.array(BO.ys), BO.model)
    acq_list = []
    i = 0
    while not stop_criteria(BO) and not BO.flag:
        if hyper_params is not None and i % 10 == 0:
            old = [math.log(get_lengthscale(BO.model)[0]), math.log(get_scale(BO.model)[0])]
            m2 = optimize_hyperparameters(BO.model, np.array(BO.xs), np.array(BO.ys), old,
                                          scale_std=float(np.ravel(sd)[0]),
                                          length_scale_only=(hyper_params == "length_scale_only"),
                                          num_restarts=num_restarts_HP, domain=BO.domain, rng=rng)
            BO.model = update_surrogate(m2, np.array(BO.xs), np.array(BO.ys))
        x_cand = optimize_acquisition(BO.acq, BO.model, BO.domain, n_grid=n_grid, n_local=n_local, rng=rng,
                                      refine=refine)
        acq_list.append(float(BO.acq(BO.model, x_cand[None, :])[0]))
        y = np.asarray(BO.func(x_cand), dtype=np.float64)
        s0 = float(np.ravel(sd)[0])
        y = y + math.sqrt(BO.noise) / s0 * rng.standard_normal(y.shape) if BO.noise > 0 else y
        BO.ys_non_std.append(y)
        y_std = (y - mu) / s0 if isinstance(BO.model, GradientGP) else (y - mu) / sd
        BO = update_bo(BO, x_cand, y_std, i)
        i += 1
    return BO, acq_list, (mu, sd)
