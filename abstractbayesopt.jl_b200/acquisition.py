"""Acquisition functions behind the reference's AbstractAcquisition API (src/abstract.jl:49):
functor `acq(surrogate, x)` over a SET of points, `update(acq, ys, surrogate)`, `copy`.
For the GPU surrogates the posterior mean, variance and the formula are evaluated in ONE fused
sweep (abo_acq_eval) instead of two posterior passes (ExpectedImprovement.jl:41-42)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .surrogates import _as_points, _get_minimum, _need_posterior
from .parallel import merge_topk

ACQ_EI, ACQ_PI, ACQ_UCB = 0, 1, 2


class AbstractAcquisition:
    acq_id = -1

    def params(self):
        raise NotImplementedError

    def __call__(self, surrogate, x):
        h = _need_posterior(surrogate)
        scores, _, _ = h.acq_eval(self.acq_id, self.params(), _as_points(x, h.d), k=0)
        return scores

    def value_and_grad(self, surrogate, x):
        """acq(surrogate, x) and its analytic gradient w.r.t. each point, one batched call."""
        h = _need_posterior(surrogate)
        return h.acq_eval_grad(self.acq_id, self.params(), _as_points(x, h.d))

    def topk(self, surrogate, x, k, want_scores=True):
        """scores and sortperm(scores; rev=true)[1:k] (acq_utils.jl:50-52) in one call; 0-based.
        want_scores=False returns None for the scores: only the k selected (index, value) pairs
        leave the device."""
        h = _need_posterior(surrogate)
        return h.acq_eval(self.acq_id, self.params(), _as_points(x, h.d), k=k, want_scores=want_scores)


@dataclass(frozen=True)
class ExpectedImprovement(AbstractAcquisition):
    """ExpectedImprovement(ξ, best_y) (ExpectedImprovement.jl:13-16)."""
    xi: float
    best_y: float
    acq_id = ACQ_EI

    def params(self): return [self.xi, self.best_y]
    def copy(self): return ExpectedImprovement(self.xi, self.best_y)
    def update(self, ys, surrogate): return ExpectedImprovement(self.xi, _get_minimum(surrogate, ys))


@dataclass(frozen=True)
class ProbabilityImprovement(AbstractAcquisition):
    """ProbabilityImprovement(ξ, best_y) (ProbabilityImprovement.jl:11-14)."""
    xi: float
    best_y: float
    acq_id = ACQ_PI

    def params(self): return [self.xi, self.best_y]
    def copy(self): return ProbabilityImprovement(self.xi, self.best_y)
    def update(self, ys, surrogate): return ProbabilityImprovement(self.xi, _get_minimum(surrogate, ys))


@dataclass(frozen=True)
class UpperConfidenceBound(AbstractAcquisition):
    """UpperConfidenceBound(β) (UpperConfidenceBound.jl:12-14)."""
    beta: float
    acq_id = ACQ_UCB

    def params(self): return [self.beta]
    def copy(self): return UpperConfidenceBound(self.beta)
    def update(self, ys, surrogate): return self


@dataclass(frozen=True)
class GradientNormUCB(AbstractAcquisition):
    """GradientNormUCB(β) (gradNormUCB.jl:12-51): UCB on the squared 2-norm of the gradient of a
    GradientGP.  Per point: m = posterior gradient mean, Σ = posterior gradient covariance (p-1 x p-1),
    -(m·m + tr Σ) + β sqrt(max(4 mᵀΣm + 2‖Σ‖_F², 1e-12)).  Means and the p x p covariance blocks come
    from batched device calls (abo_gp_posterior / abo_gp_posterior_cov), chunked so that
    points*p <= 4096; there is no fused sweep for this one."""
    beta: float

    def params(self): return [self.beta]
    def copy(self): return GradientNormUCB(self.beta)
    def update(self, ys, surrogate): return self

    def __call__(self, surrogate, x):
        h = _need_posterior(surrogate)
        X = _as_points(x, h.d)
        p = h.p
        if p < 2:
            raise TypeError("GradientNormUCB needs a GradientGP surrogate")
        out = np.empty(len(X))
        step = max(1, 4096 // p)
        for c0 in range(0, len(X), step):
            Xc = X[c0:c0 + step]
            mc = len(Xc)
            mean, _ = h.posterior(Xc, p, True, False)
            cov = h.posterior_cov(Xc, p)
            mean = mean.reshape(p, mc)
            for c in range(mc):
                idx = np.arange(1, p) * mc + c
                m = mean[1:, c]
                S = cov[np.ix_(idx, idx)]
                mu_sq = float(m @ m + np.trace(S))
                var_sq = float(4.0 * m @ (S @ m) + 2.0 * np.sum(S * S))
                out[c0 + c] = -mu_sq + self.beta * np.sqrt(max(var_sq, 1e-12))
        return out

    def topk(self, surrogate, x, k, want_scores=True):
        s = self(surrogate, x)
        ti, tv = merge_topk([np.arange(len(s))], [s], k)
        return s, ti, tv


class EnsembleAcquisition(AbstractAcquisition):
    """EnsembleAcquisition(weights, acquisitions) (EnsembleAcq.jl:12-55): non-negative weights,
    normalised to sum 1; value = Σ w_i acq_i(surrogate, x)."""

    def __init__(self, weights, acquisitions):
        w = np.asarray(weights, dtype=np.float64)
        if len(w) != len(acquisitions):
            raise ValueError("weights and acquisitions must align")
        if np.any(w < 0):
            raise ValueError("weights must be non-negative")
        if not w.sum() > 0:
            raise ValueError("sum of weights must be positive")
        self.weights = w / w.sum()
        self.acquisitions = list(acquisitions)

    def copy(self): return EnsembleAcquisition(self.weights.copy(), [a.copy() for a in self.acquisitions])
    def update(self, ys, surrogate): return EnsembleAcquisition(self.weights, [a.update(ys, surrogate) for a in self.acquisitions])

    def __call__(self, surrogate, x):
        return sum(w * a(surrogate, x) for w, a in zip(self.weights, self.acquisitions))

    def topk(self, surrogate, x, k, want_scores=True):
        s = self(surrogate, x)
        ti, tv = merge_topk([np.arange(len(s))], [s], k)
        return s, ti, tv
