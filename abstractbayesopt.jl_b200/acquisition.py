"""Acquisition functions behind the reference's AbstractAcquisition API (src/abstract.jl:49):
functor `acq(surrogate, x)` over a SET of points, `update(acq, ys, surrogate)`, `copy`.
For the GPU surrogates the posterior mean, variance and the formula are evaluated in ONE fused
sweep (abo_acq_eval) instead of two posterior passes (ExpectedImprovement.jl:41-42)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .surrogates import _as_points, _get_minimum, _need_posterior

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

    def topk(self, surrogate, x, k):
        """scores and sortperm(scores; rev=true)[1:k] (acq_utils.jl:50-52) in one call; 0-based."""
        h = _need_posterior(surrogate)
        return h.acq_eval(self.acq_id, self.params(), _as_points(x, h.d), k=k)


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
