"""Lanczos quadrature for spectral densities of symmetric linear operators
(``vivit/hessianfree/lanczos.py:13-270``; Algorithm 2 and Section D.2 of Papyan, "Traces of class/cross-class
structure pervade deep learning spectra", 2020).

Any SciPy ``LinearOperator`` works.  Operators of ``vivit_b200.hessianfree`` additionally expose
``matvec_torch``; for them the three-term recurrence runs on the operator's device without a host copy per
iteration, and on a GPU the tridiagonal matrix is decomposed by the library's own eigensolver (``vvt_syevj``).
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
from scipy.linalg import eigh_tridiagonal
from scipy.sparse.linalg import LinearOperator, eigsh


def _recurrence_host(A: LinearOperator, ncv: int) -> Tuple[np.ndarray, np.ndarray]:
    alphas, betas = np.zeros(ncv), np.zeros(max(ncv - 1, 0))
    v = np.random.randn(A.shape[1])
    v /= np.linalg.norm(v)
    v_prev, beta = None, 0.0
    for m in range(ncv):
        w = A @ v
        if m > 0:
            w = w - beta * v_prev
        alphas[m] = np.inner(w, v)
        w = w - alphas[m] * v
        if m < ncv - 1:
            beta = betas[m] = np.linalg.norm(w)
            v_prev, v = v, w / beta
    return alphas, betas


def _recurrence_device(A, ncv: int):
    """Same recurrence on torch tensors living where the operator lives; returns the coefficient tensors."""
    import torch

    device = A._device
    v = torch.from_numpy(np.random.randn(A.shape[1])).to(device)
    v = (v / v.norm()).to(A._params[0].dtype)
    alphas = torch.zeros(ncv, dtype=torch.float64, device=device)
    betas = torch.zeros(max(ncv - 1, 0), dtype=torch.float64, device=device)
    v_prev, beta = None, None
    for m in range(ncv):
        w = A.matvec_torch(v)
        if m > 0:
            w = w - beta * v_prev
        alpha = torch.dot(w, v)
        alphas[m] = alpha
        w = w - alpha * v
        if m < ncv - 1:
            beta = w.norm()
            betas[m] = beta
            v_prev, v = v, w / beta
    return alphas, betas


def _eigh_tridiagonal_device(alphas, betas) -> Tuple[np.ndarray, np.ndarray]:
    """Eigen-decomposition of the Lanczos matrix on the GPU (dense ``ncv x ncv``, float64, ``vvt_syevj``)."""
    import torch

    from vivit_b200 import kernels

    T = torch.diag(alphas)
    radius = torch.zeros_like(alphas)
    if betas.numel():
        T = T + torch.diag(betas, 1) + torch.diag(betas, -1)
        radius[:-1] += betas.abs()
        radius[1:] += betas.abs()
    # vvt_syevj is written for Gram matrices (positive semi-definite; it factors G + eps I).  A Hessian's
    # Lanczos matrix is indefinite: shift it by its Gershgorin lower bound, which moves every eigenvalue by
    # exactly that amount and leaves the eigenvectors alone.
    shift = torch.clamp(-(alphas - radius).min(), min=0.0)
    T = T + shift * torch.eye(T.shape[0], dtype=T.dtype, device=T.device)
    evals, evecs = kernels.syevj(T.contiguous(), vectors=True)
    return (evals - shift).cpu().numpy(), evecs.cpu().numpy()


def fast_lanczos(A: LinearOperator, ncv: int, use_eigh_tridiagonal: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """``ncv`` Lanczos iterations without re-orthogonalisation (``lanczos.py:13-62``).

    Returns the eigenvalues (ascending) and eigenvectors of the tridiagonal matrix; ``evecs[:, i]`` is the
    normalised eigenvector of ``evals[i]`` and ``evecs[0, i] ** 2`` the quadrature weight of that node.
    """
    if hasattr(A, "matvec_torch"):
        alphas_t, betas_t = _recurrence_device(A, ncv)
        if alphas_t.is_cuda and not use_eigh_tridiagonal:
            return _eigh_tridiagonal_device(alphas_t, betas_t)
        alphas, betas = alphas_t.cpu().numpy(), betas_t.cpu().numpy()
    else:
        alphas, betas = _recurrence_host(A, ncv)
    if use_eigh_tridiagonal:
        return eigh_tridiagonal(alphas, betas)
    T = np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1)
    return np.linalg.eigh(T)


def approximate_boundaries(A: LinearOperator, tol: float = 1e-2) -> Tuple[float, float]:
    """Estimates of the smallest and largest eigenvalue of ``A`` (ARPACK, ``lanczos.py:65-79``)."""
    eval_min, eval_max = eigsh(A, k=2, which="BE", tol=tol, return_eigenvectors=False)
    return eval_min, eval_max


def approximate_boundaries_abs(A: LinearOperator, tol: float = 1e-2) -> Tuple[float, float]:
    """Estimates of the smallest and largest eigenvalue of ``|A|`` (``lanczos.py:82-99``)."""
    (eval_max,) = eigsh(A, k=1, which="LM", tol=tol, return_eigenvectors=False)
    (eval_min,) = eigsh(A, k=1, which="SM", tol=tol, return_eigenvectors=False)
    return abs(eval_min), abs(eval_max)


def _bumps(grid: np.ndarray, nodes: np.ndarray, weights: np.ndarray, sigma: float) -> np.ndarray:
    """``sum_i weights_i N(grid; nodes_i, sigma)``."""
    z = (grid[None, :] - nodes[:, None]) / sigma
    return (weights[:, None] * np.exp(-0.5 * z * z)).sum(0) / (sigma * np.sqrt(2 * np.pi))


def _interval(lo: float, hi: float, margin: float) -> Tuple[float, float]:
    """Centre and half-width of ``[lo, hi]`` widened by ``margin`` on both sides."""
    pad = margin * (hi - lo)
    lo, hi = lo - pad, hi + pad
    return (hi + lo) / 2, (hi - lo) / 2


def lanczos_approximate_spectrum(
    A: LinearOperator,
    ncv: int,
    num_points: int = 1024,
    num_repeats: int = 1,
    kappa: float = 3.0,
    boundaries: Optional[Tuple[float, float]] = None,
    margin: float = 0.05,
    boundaries_tol: float = 1e-2,
) -> Tuple[np.ndarray, np.ndarray]:
    """Spectral density ``p(l) = 1/d sum_i delta(l - l_i)`` of ``A`` on a grid (``lanczos.py:102-170``).

    The spectrum is mapped to ``[-1, 1]`` (``(A - c I) / d``) so that the width of the Gaussian bumps that
    replace the delta peaks, set by ``kappa > 1``, needs no tuning; ``num_repeats`` quadratures are averaged.
    Returns the grid points and the density.
    """
    if boundaries is None:
        boundaries = approximate_boundaries(A, tol=boundaries_tol)
    c, d = _interval(boundaries[0], boundaries[1], margin)
    grid_norm = np.linspace(-1, 1, num_points, endpoint=True)
    sigma = 2 / (ncv - 1) / np.sqrt(8 * np.log(kappa))
    density = np.zeros_like(grid_norm)
    for _ in range(num_repeats):
        evals, evecs = fast_lanczos(A, ncv)
        density += _bumps(grid_norm, (evals - c) / d, evecs[0, :] ** 2 / d, sigma) / num_repeats
    return grid_norm * d + c, density


def lanczos_approximate_log_spectrum(
    A: LinearOperator,
    ncv: int,
    num_points: int = 1024,
    num_repeats: int = 1,
    kappa: float = 1.04,
    boundaries: Optional[Tuple[float, float]] = None,
    margin: float = 0.05,
    boundaries_tol: float = 1e-2,
    epsilon: float = 1e-5,
) -> Tuple[np.ndarray, np.ndarray]:
    """Spectral density of ``log(|A| + epsilon I)`` (natural log), returned on the grid of ``|l| + epsilon``
    values (``lanczos.py:187-270``).  ``boundaries`` are estimates of the extreme eigenvalues of ``|A|``."""
    if boundaries is None:
        boundaries = approximate_boundaries_abs(A, tol=boundaries_tol)
    c, d = _interval(np.log(boundaries[0] + epsilon), np.log(boundaries[1] + epsilon), margin)
    grid_norm = np.linspace(-1, 1, num_points, endpoint=True)
    grid_out = np.exp(grid_norm * d + c)
    sigma = 2 / (ncv - 1) / np.sqrt(8 * np.log(kappa))
    density = np.zeros_like(grid_norm)
    for _ in range(num_repeats):
        evals, evecs = fast_lanczos(A, ncv)
        nodes = (np.log(np.abs(evals) + epsilon) - c) / d
        density += _bumps(grid_norm, nodes, evecs[0, :] ** 2, sigma) / num_repeats
    return grid_out, density / (d * grid_out)
