"""ContinuousDomain (src/domains/ContinuousDomain.jl:16-28): a box with validated bounds."""
from __future__ import annotations

import numpy as np


class AbstractDomain:
    pass


class ContinuousDomain(AbstractDomain):
    def __init__(self, lower, upper):
        lower = np.asarray(lower, dtype=np.float64).reshape(-1)
        upper = np.asarray(upper, dtype=np.float64).reshape(-1)
        if lower.size != upper.size:
            raise ValueError("lower and upper must have the same length")
        if not np.all(lower <= upper):
            raise ValueError("lower bounds must be less than or equal to upper bounds")
        self.lower, self.upper = lower, upper
        self.bounds = [(float(a), float(b)) for a, b in zip(lower, upper)]      # ContinuousDomain.jl: Vector{Tuple}
