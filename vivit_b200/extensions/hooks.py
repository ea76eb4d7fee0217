"""Gram-matrix extension hooks, API of ``vivit/extensions/hooks.py``.

* ``GramBatchGrad`` / ``CenteredGramBatchGrad`` / ``CenteredBatchGrad``
  (``vivit/extensions/firstorder/batch_grad/gram_batch_grad.py:7-213``): used as
  ``extension_hook`` next to ``BatchGrad``; ``[N, N]`` Gram matrix of the (centred)
  per-sample gradients, summed over parameters.
* ``GramSqrtGGNExact`` / ``GramSqrtGGNMC``
  (``vivit/extensions/secondorder/sqrt_ggn/gram_sqrt_ggn.py:9-142``): used next to
  ``SqrtGGNExact`` / ``SqrtGGNMC``; ``[C N, C N]`` Gram matrix of the GGN factor.

The savefields written by ``BatchGrad`` / ``SqrtGGN*`` hold either the materialised
tensors (the extensions' default, as in [BackPACK]) or, with ``lazy=True``, the
structured ``Factor`` / ``GradFactor`` objects; both are accepted.  All arithmetic runs
in the CUDA kernels behind ``vivit_b200.kernels`` (dense Gram on tcgen05 / DMMA, structured
Linear Gram, ``vvt_center_rows``).  Like the reference's, a hook object is single-use.
"""

from __future__ import annotations

import torch
from torch import Tensor

from vivit_b200 import kernels
from vivit_b200.factors import Factor, GradFactor
from vivit_b200.utils.hooks import ParameterHook

__all__ = [
    "GramBatchGrad",
    "CenteredBatchGrad",
    "CenteredGramBatchGrad",
    "GramSqrtGGNExact",
    "GramSqrtGGNMC",
]


def _gram_of_rows(rows: Tensor) -> Tensor:
    """``rows [R, D] -> rows rows^T`` (``pairwise_dot`` + ``reshape_as_square``,
    ``vivit/utils/gram.py:9-35,58-69``)."""
    rows = rows.detach()
    G = torch.zeros(rows.shape[0], rows.shape[0], dtype=rows.dtype, device=rows.device)
    return kernels.gram_dense_accum(G, rows)


class CenteredBatchGrad(ParameterHook):
    """Store ``grad_batch - grad_batch.mean(0)`` (``[N, *param.shape]``) under ``savefield``
    (``gram_batch_grad.py:7-38``).  ``grad_batch`` itself is left untouched."""

    _SAVEFIELD_GRAD_BATCH = "grad_batch"

    def __init__(self, savefield: str = "centered_grad_batch"):
        super().__init__(savefield)

    def param_hook(self, param):
        grad_batch = getattr(param, self._SAVEFIELD_GRAD_BATCH)
        if isinstance(grad_batch, GradFactor):
            grad_batch = grad_batch.materialize()
        return kernels.center_rows(grad_batch.detach())


class _GramBatchGradBase(ParameterHook):
    """``gram_batch_grad.py:41-123``: per parameter, (optionally centre in place and) add the
    pairwise dot products of the per-sample gradients to the running ``[N, N]`` result."""

    _SAVEFIELD_GRAD_BATCH = "grad_batch"

    def __init__(self, savefield, center, layerwise=False, free_grad_batch=False):
        super().__init__(savefield)
        self._center = center
        self._gram_mat = None
        self._layerwise = layerwise
        self._free_grad_batch = free_grad_batch

    def param_hook(self, param):
        grad_batch = getattr(param, self._SAVEFIELD_GRAD_BATCH)
        if isinstance(grad_batch, GradFactor):
            rows = grad_batch.dense()
            if self._center:  # a structured factor has no buffer to centre in place
                rows = kernels.center_rows(rows)
        else:
            if self._center:
                # the reference centres the stored tensor itself (`grad_batch -= mean`, :88-89)
                if grad_batch.is_contiguous():
                    kernels.center_rows(grad_batch, inplace=True)
                else:
                    grad_batch.copy_(kernels.center_rows(grad_batch))
            rows = grad_batch.reshape(grad_batch.shape[0], -1)
        gram_param = _gram_of_rows(rows)
        self._update_result(gram_param)
        if self._free_grad_batch:
            delattr(param, self._SAVEFIELD_GRAD_BATCH)
        if self._layerwise:
            return gram_param

    def get_result(self):
        """``[N, N]`` Gram matrix of the per-sample gradients of all parameters."""
        return self._gram_mat

    def _update_result(self, mat):
        if self._gram_mat is None:
            # the first parameter's matrix may also be handed out layer-wise: keep them apart
            self._gram_mat = mat.clone() if self._layerwise else mat
        else:
            self._gram_mat += mat


class GramBatchGrad(_GramBatchGradBase):
    """Uncentred gradient Gram matrix ``<g_i, g_j>`` (``gram_batch_grad.py:126-168``); with a
    mean-reduced loss the ``g_i`` carry the ``1/N`` of [BackPACK]'s ``BatchGrad``."""

    def __init__(self, savefield="gram_grad_batch", layerwise=False, free_grad_batch=False):
        super().__init__(savefield, False, layerwise=layerwise, free_grad_batch=free_grad_batch)


class CenteredGramBatchGrad(_GramBatchGradBase):
    """Centred gradient Gram matrix ``<g_i - mean g, g_j - mean g>`` (``gram_batch_grad.py:171-213``)."""

    def __init__(self, savefield="centered_gram_grad_batch", layerwise=False, free_grad_batch=False):
        super().__init__(savefield, True, layerwise=layerwise, free_grad_batch=free_grad_batch)


class GramSqrtGGN(ParameterHook):
    """``gram_sqrt_ggn.py:9-74``: add ``pairwise_dot(sqrt_ggn, start_dim=2)`` of every parameter
    to the running ``[C N, C N]`` result."""

    SQRT_GGN_SAVEFIELDS = {"exact": "sqrt_ggn_exact", "sampling": "sqrt_ggn_mc"}

    def __init__(self, loss_hessian_strategy, savefield, layerwise, free_sqrt_ggn):
        super().__init__(savefield)
        self._gram_mat = None
        self._layerwise = layerwise
        self._free_sqrt_ggn = free_sqrt_ggn
        self._savefield_sqrt_ggn = self.SQRT_GGN_SAVEFIELDS[loss_hessian_strategy]

    def param_hook(self, param):
        sqrt_ggn = getattr(param, self._savefield_sqrt_ggn)
        if isinstance(sqrt_ggn, Factor):
            gram_param = sqrt_ggn.gram_mat().reshape(sqrt_ggn.R, sqrt_ggn.R)
        else:
            R = sqrt_ggn.shape[0] * sqrt_ggn.shape[1]
            gram_param = _gram_of_rows(sqrt_ggn.reshape(R, -1))
        self._update_result(gram_param)
        if self._free_sqrt_ggn:
            delattr(param, self._savefield_sqrt_ggn)
        if self._layerwise:
            return gram_param

    def get_result(self):
        """``[C N, C N]`` (``[M N, M N]``) GGN Gram matrix of all parameters."""
        return self._gram_mat

    def _update_result(self, mat):
        if self._gram_mat is None:
            self._gram_mat = mat.clone() if self._layerwise else mat
        else:
            self._gram_mat += mat


class GramSqrtGGNExact(GramSqrtGGN):
    """GGN Gram matrix from ``SqrtGGNExact`` (``gram_sqrt_ggn.py:77-107``)."""

    def __init__(self, savefield="gram_sqrt_ggn_exact", layerwise=False, free_sqrt_ggn=False):
        super().__init__("exact", savefield, layerwise, free_sqrt_ggn)


class GramSqrtGGNMC(GramSqrtGGN):
    """GGN Gram matrix from ``SqrtGGNMC`` (``gram_sqrt_ggn.py:110-142``)."""

    def __init__(self, savefield="gram_sqrt_ggn_mc", layerwise=False, free_sqrt_ggn=False):
        super().__init__("sampling", savefield, layerwise, free_sqrt_ggn)
