"""Symmetric eigendecompositions with the reference's conveniences (``vivit/utils/eig.py``), on ``vvt_syevj``.

* ``symeig_psd``: decomposition of ``input + shift * I`` with the shift taken off the eigenvalues again
  (``utils/eig.py:6-48``).  The reference needs the shift to get LAPACK through ill-conditioned PSD
  matrices; ``vvt_syevj`` converges on those without it (``tests/golden/symeig_killer.pt``), the argument is
  kept for drop-in use.
* ``symeig``: decomposition followed by ``remove_zero_evals`` (``utils/eig.py:77-108``).
* ``remove_zero_evals``: drops pairs whose eigenvalue ``isclose`` to zero (``utils/eig.py:111-134``;
  ``vvt_filter_nonzero``).
* ``shift_diag``: ``utils/eig.py:51-74``.

Conventions of ``Tensor.symeig``: ascending eigenvalues, eigenvectors as columns, the ``upper`` (default) or
lower triangle is read, and an EMPTY tensor stands for the eigenvectors when they were not asked for.  The
decomposition and the filter run in the CUDA library; the diagonal shift and ``evals - shift`` are ``O(R)``
device-side glue.
"""

from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from vivit_b200 import kernels


def shift_diag(input: Tensor, shift: float, inplace: bool = False) -> Tensor:
    """``input`` with ``shift`` added to its main diagonal (also of a rectangular matrix)."""
    if shift == 0.0:
        return input
    result = input if inplace else input.clone()
    result.diagonal().add_(shift)
    return result


def _decompose(matrix: Tensor, eigenvectors: bool, upper: bool) -> Tuple[Tensor, Tensor]:
    if matrix.shape[0] != matrix.shape[1]:
        raise ValueError(f"Input must be square. Got {tuple(matrix.shape)}.")
    if torch.isnan(matrix).any():
        raise RuntimeError("Tensor contains NaNs: True")
    # vvt_syevj reads the upper triangle; the lower one of A is the upper one of A^T
    matrix = matrix if upper else matrix.t().contiguous()
    # vvt_syevj is a one-sided method for positive semi-definite matrices (Grams); the reference's
    # ``Tensor.symeig`` takes any symmetric matrix.  A Gershgorin lower bound of the spectrum decides: when
    # it is negative the decomposition runs on ``A - bound * I`` (positive semi-definite by construction, same
    # eigenvectors) and the shift is taken off the eigenvalues again.
    tri = torch.triu(matrix)
    radius = (tri.abs().sum(0) + tri.abs().sum(1)) - 2 * tri.diagonal().abs()
    bound = float((tri.diagonal() - radius).min()) if matrix.numel() else 0.0
    gershgorin = -bound * (1.0 + 1e-6) if bound < 0.0 else 0.0
    if gershgorin:
        matrix = shift_diag(matrix, gershgorin)
    try:
        evals, evecs = kernels.syevj(matrix, vectors=eigenvectors)
    except kernels.SyevjNotConverged as e:
        raise RuntimeError(f"{e}. Tensor contains NaNs: False") from e
    if gershgorin:
        evals = evals - gershgorin
    if evecs is None:
        evecs = torch.empty(0, dtype=matrix.dtype, device=matrix.device)
    return evals, evecs


def symeig_psd(
    input: Tensor, eigenvectors: bool = False, upper: bool = True, shift: float = 0.0, shift_inplace: bool = False
) -> Tuple[Tensor, Tensor]:
    """Eigenvalues (ascending) and eigenvectors (columns) of a symmetric PSD matrix, computed on the matrix
    with its diagonal shifted by ``shift`` and reported for the unshifted one."""
    if input.dim() != 2:
        raise ValueError(f"Input must have dimension 2. Got {input.dim()}.")
    shifted = shift_diag(input, shift, inplace=shift_inplace)
    try:
        evals, evecs = _decompose(shifted, eigenvectors, upper)
    finally:
        if shift_inplace:
            shift_diag(input, -shift, inplace=True)
    if shift != 0.0:
        evals.sub_(shift)
    return evals, evecs


def remove_zero_evals(evals: Tensor, evecs: Tensor, atol: float = 1e-7, rtol: float = 1e-5) -> Tuple[Tensor, Tensor]:
    """Keep the pairs whose eigenvalue is not ``isclose(0, rtol, atol)``."""
    if evals.numel() == 0:
        return evals, evecs
    nonzero = kernels.filter_nonzero(evals, atol=atol, rtol=rtol)
    evals = evals[nonzero]
    if evecs.numel() != 0:
        evecs = evecs[:, nonzero]
    return evals, evecs


def symeig(
    input: Tensor, eigenvectors: bool = False, upper: bool = True, atol: float = 1e-7, rtol: float = 1e-5
) -> Tuple[Tensor, Tensor]:
    """Decomposition of a symmetric matrix without the numerically-zero eigenpairs."""
    if input.dim() != 2:
        raise ValueError("Input must be of dimension 2")
    evals, evecs = _decompose(input, eigenvectors, upper)
    return remove_zero_evals(evals, evecs, atol=atol, rtol=rtol)
