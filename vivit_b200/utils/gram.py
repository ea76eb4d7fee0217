"""Gram-matrix helpers of the reference (``vivit/utils/gram.py``) on the CUDA library.

The reference spells every one of these as an ``einsum`` over lettered indices.  Here the leading
("which vector") axes and the trailing (contracted) axes are flattened to a matrix and the contraction is
one GEMM of the C ABI: the symmetric tcgen05 / DMMA Gram kernel for ``pairwise_dot`` and
``compute_gram_mat`` (``vvt_gram_dense_accum``), ``vvt_gemm`` for the rest.
"""

from __future__ import annotations

import math
from typing import Iterable, List, Sequence, Tuple, Union

import torch
from torch import Tensor

from vivit_b200 import kernels


def _as_rows(t: Tensor, start_dim: int) -> Tuple[Tensor, Tuple[int, ...]]:
    """``[*lead, *features] -> ([prod(lead), prod(features)], lead)``."""
    lead = tuple(t.shape[:start_dim])
    return t.reshape(math.prod(lead), -1).contiguous(), lead


def reshape_as_square(tensor: Tensor) -> Tensor:
    """View a tensor whose number of entries is a perfect square as that square (``utils/gram.py:58-69``)."""
    side = math.isqrt(tensor.numel())
    if side * side != tensor.numel():
        raise ValueError(f"{tensor.numel()} entries do not form a square matrix")
    return tensor.reshape(side, side)


def partial_contract(tensor: Tensor, other: Tensor, start_dims: Sequence[int]) -> Tensor:
    """All scalar products between the slices ``tensor[i...]`` and ``other[j...]``; the slices start at axis
    ``start_dims[0]`` / ``start_dims[1]`` (``utils/gram.py:206-232``).  Shape ``[*lead(tensor), *lead(other)]``."""
    d1, d2 = start_dims
    if tensor.dim() - d1 != other.dim() - d2:
        raise ValueError("Trailing dimensions don't match.")
    a, lead_a = _as_rows(tensor, d1)
    b, lead_b = _as_rows(other, d2)
    return kernels.gemm(a, b).reshape(*lead_a, *lead_b)


def pairwise_dot(tensor: Tensor, start_dim: int = 1, flatten: bool = True) -> Tensor:
    """Gram matrix of the slices of ``tensor`` that start at axis ``start_dim`` (``utils/gram.py:9-35``):
    square ``[A, A]`` if ``flatten`` else ``[*lead, *lead]``."""
    rows, lead = _as_rows(tensor, start_dim)
    gram = torch.zeros(rows.shape[0], rows.shape[0], dtype=rows.dtype, device=rows.device)
    kernels.gram_dense_accum(gram, rows)
    return gram if flatten else gram.reshape(*lead, *lead)


def compute_gram_mat(parameters: Iterable, savefield: str, start_dim: int, flatten: bool = True):
    """Sum over ``parameters`` of ``pairwise_dot(getattr(p, savefield), start_dim)``
    (``utils/gram.py:72-116``); accumulated in place by the Gram kernel.  ``None`` without parameters."""
    gram, lead = None, ()
    for p in parameters:
        rows, lead = _as_rows(getattr(p, savefield), start_dim)
        if gram is None:
            gram = torch.zeros(rows.shape[0], rows.shape[0], dtype=rows.dtype, device=rows.device)
        kernels.gram_dense_accum(gram, rows)
    if gram is None or flatten:
        return gram
    return gram.reshape(*lead, *lead)


def sqrt_gram_mat_prod(
    mat: Tensor, parameters: Iterable, savefield: str, start_dim: int, concat: bool = False
) -> Union[List[Tensor], Tensor]:
    """``U @ mat`` per parameter, ``U = getattr(p, savefield)`` flattened to ``[A, *p.shape]`` and read as
    ``[D_p, A]`` (the Gram matrix is ``U^T U``): results ``[*p.shape, J]``, or one ``[D, J]`` matrix if
    ``concat`` (``utils/gram.py:119-179``)."""
    if mat.dim() != 2:
        raise NotImplementedError("Can only multiply with matrices")
    result = []
    for p in parameters:
        rows, _ = _as_rows(getattr(p, savefield), start_dim)  # [A, D_p]
        prod = kernels.gemm(rows, mat, trans_a=True, trans_b=True)  # [D_p, J]
        result.append(prod.reshape(*getattr(p, savefield).shape[start_dim:], mat.shape[1]))
    if concat:
        return torch.cat([r.reshape(-1, mat.shape[1]) for r in result])
    return result


def mVp(V_t: Tensor, mat: Tensor, start_dim: int) -> Tensor:
    """``V^T`` applied to stacked parameter-shaped vectors: ``mat [F, *p.shape]``, ``V_t [*lead, *p.shape]``
    -> ``[F, *lead]`` (``utils/gram.py:182-203``)."""
    rows, lead = _as_rows(V_t, start_dim)
    return kernels.gemm(mat.reshape(mat.shape[0], -1).contiguous(), rows).reshape(mat.shape[0], *lead)


def split_list(sequence: Sequence, lengths: Sequence[int]):
    """Consecutive sub-lists of the given lengths (``utils/gram.py:235-256``)."""
    if len(sequence) != sum(lengths):
        raise ValueError("Sub-list lengths don't sum to length of the full list.")
    start = 0
    for length in lengths:
        yield sequence[start : start + length]
        start += length
