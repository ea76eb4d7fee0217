"""Per-sample directional derivatives along the leading GGN eigenvectors, and the damped Newton step
assembled from them."""

from vivit_b200.optim import directional_damped_newton as _newton
from vivit_b200.optim import directional_derivatives as _derivatives

DirectionalDerivativesComputation = _derivatives.DirectionalDerivativesComputation
DirectionalDampedNewtonComputation = _newton.DirectionalDampedNewtonComputation

__all__ = ["DirectionalDerivativesComputation", "DirectionalDampedNewtonComputation"]
