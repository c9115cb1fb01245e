"""Kernel objects mirroring the KernelFunctions / in-repo kernels the reference composes as
`σ² * with_lengthscale(base, ℓ)`  ==  ScaledKernel(TransformedKernel(base, ScaleTransform(1/ℓ)), σ²)
(src/surrogates/StandardGP.jl:41-64, surrogates_utils.jl:28-47)."""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

KERNEL_IDS = {
    "SqExponentialKernel": 0, "Matern52Kernel": 1, "Matern72Kernel": 2,
    "ApproxMatern52Kernel": 3, "ApproxMatern72Kernel": 4, "ADMatern52Kernel": 5, "ADMatern72Kernel": 6,
}


@dataclass(frozen=True)
class Kernel:
    name: str
    inv_lengthscale: float | tuple | None = None   # ScaleTransform s = 1/ℓ (None: no transform yet); a tuple = ARDTransform, s_k = 1/ℓ_k
    scale: float | None = None             # ScaledKernel σ² (None: not scaled yet)

    @property
    def kernel_id(self) -> int:
        return KERNEL_IDS[self.name]

    def __rmul__(self, sigma2):            # σ² * kernel
        base = 1.0 if self.scale is None else self.scale
        return replace(self, scale=float(sigma2) * base)

    def constructor(self) -> "Kernel":     # get_kernel_constructor: the bare base kernel
        return Kernel(self.name)


def SqExponentialKernel(): return Kernel("SqExponentialKernel")
def Matern52Kernel(): return Kernel("Matern52Kernel")
def Matern72Kernel(): return Kernel("Matern72Kernel")
def ApproxMatern52Kernel(): return Kernel("ApproxMatern52Kernel")
def ApproxMatern72Kernel(): return Kernel("ApproxMatern72Kernel")
def ADMatern52Kernel(): return Kernel("ADMatern52Kernel")
def ADMatern72Kernel(): return Kernel("ADMatern72Kernel")


def with_lengthscale(kernel: Kernel, lengthscale) -> Kernel:
    """KernelFunctions.with_lengthscale: base ∘ ScaleTransform(1/ℓ); a vector of length scales gives
    base ∘ ARDTransform(1 ./ ℓ) (one length scale per input dimension)."""
    if np.ndim(lengthscale) > 0:
        return replace(kernel, inv_lengthscale=tuple(1.0 / float(v) for v in np.ravel(lengthscale)))
    return replace(kernel, inv_lengthscale=1.0 / float(lengthscale))


def extract_scale_and_lengthscale(kernel: Kernel):
    """src/surrogates/surrogates_utils.jl:28-47 — returns (inner, scale, lengthscale|None).
    The lengthscale is 1/s, and the constructor re-applies with_lengthscale(inner, 1/s), i.e.
    the stored s goes through 1/(1/s) exactly as in the reference."""
    scale = 1.0 if kernel.scale is None else kernel.scale
    if isinstance(kernel.inv_lengthscale, tuple):
        ls = tuple(1.0 / v for v in kernel.inv_lengthscale)
    else:
        ls = None if kernel.inv_lengthscale is None else 1.0 / kernel.inv_lengthscale
    return kernel.constructor(), scale, ls


def normalized(kernel: Kernel) -> Kernel:
    """What StandardGP(kernel, …) stores: always Scaled(Transformed(base)) (StandardGP.jl:47-59)."""
    if kernel.inv_lengthscale is not None and kernel.scale is not None:
        return kernel                      # already Scaled(Transformed(base)): the prior is shared, not rebuilt
    inner, scale, ls = extract_scale_and_lengthscale(kernel)
    inner = with_lengthscale(inner, 1.0 if ls is None else ls)
    return replace(inner, scale=scale)
