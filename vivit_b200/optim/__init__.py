"""Directional derivatives and damped Newton steps along GGN eigenvectors."""

from vivit_b200.optim.directional_damped_newton import DirectionalDampedNewtonComputation
from vivit_b200.optim.directional_derivatives import DirectionalDerivativesComputation

__all__ = ["DirectionalDampedNewtonComputation", "DirectionalDerivativesComputation"]
