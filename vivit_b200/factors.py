"""Per-parameter representations of the GGN factor ``V_p^T`` and of per-sample gradients.

The reference materialises ``V_p^T`` as a ``[C, N, *param.shape]`` tensor for every
parameter except 2-d ``Linear.weight`` (``vivit/extensions/secondorder/vivit/base.py:84-92``
vs ``linear.py:41-81``), and its optim classes materialise it for *all* parameters
(``vivit/optim/directional_derivatives.py:238-247``).  Here a factor is an object
that knows how to contribute to the Gram matrix, to the cross term with the
per-sample gradients, and how to apply ``V_p`` -- so Linear layers stay
structured (``S`` and ``Z`` only) on every path and nothing of size
``R x out x in`` is ever written.

All heavy lifting is in ``vivit_b200.kernels`` (hand-written CUDA via the C ABI).
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from vivit_b200 import kernels


class Factor:
    """``V_p^T`` of one parameter: ``C`` classes (or MC samples) x ``N`` samples rows."""

    C: int
    N: int
    param_shape: Tuple[int, ...]

    @property
    def R(self) -> int:
        return self.C * self.N

    # G [R, R] += V_p^T V_p   (fold: honour the weight/bias folding of a Linear layer, see LinearBiasFactor)
    def gram_accum(self, G: Tensor, fold: bool = True) -> None:
        raise NotImplementedError

    # X [R, n_g] += V_p^T g_p^T
    def cross_accum(self, X: Tensor, grad: "GradFactor") -> None:
        raise NotImplementedError

    # U [K, R] -> [K, *param_shape]; adds squared norms into norm2 (float64 [K]) if given
    def backtransform(self, U: Tensor, norm2: Optional[Tensor]) -> Tensor:
        raise NotImplementedError

    # v [R] -> param_shape
    def v_apply(self, v: Tensor) -> Tensor:
        raise NotImplementedError

    # M [F, *param_shape] -> [F, C, N]
    def vt_mat_prod(self, M: Tensor) -> Tensor:
        raise NotImplementedError

    # [C, N, *param_shape]
    def materialize(self) -> Tensor:
        raise NotImplementedError

    def gram_mat(self) -> Tensor:
        """``[C, N, C, N]`` Gram matrix of this parameter alone (``base.py:118-124``)."""
        like = self._like()
        G = torch.zeros(self.R, self.R, dtype=like.dtype, device=like.device)
        self.gram_accum(G, fold=False)
        return G.reshape(self.C, self.N, self.C, self.N)

    def _like(self) -> Tensor:
        raise NotImplementedError


class DenseFactor(Factor):
    """Materialised ``V_p^T`` in the coalesced ``[R, D_p]`` layout (conv weights/biases,
    Linear bias, Linear with extra input dimensions)."""

    def __init__(self, Vt: Tensor, param_shape):
        # Vt: [C, N, *param_shape]
        self.C, self.N = Vt.shape[:2]
        self.param_shape = tuple(param_shape)
        self.Vt = Vt.reshape(self.C * self.N, -1)

    def _like(self):
        return self.Vt

    def gram_accum(self, G, fold=True):
        kernels.gram_dense_accum(G, self.Vt)

    def cross_accum(self, X, grad):
        kernels.gram_cross_accum(X, self.Vt, grad.dense())

    def backtransform(self, U, norm2):
        E = kernels.backtransform_dense(U.reshape(U.shape[0], -1), self.Vt, norm2)
        return E.reshape(U.shape[0], *self.param_shape)

    def v_apply(self, v):
        return kernels.v_apply_dense(v.reshape(-1), self.Vt).reshape(self.param_shape)

    def vt_mat_prod(self, M):
        F_ = M.shape[0]
        out = torch.zeros(self.R, F_, dtype=self.Vt.dtype, device=self.Vt.device)
        kernels.gram_cross_accum(out, self.Vt, M.reshape(F_, -1))
        return out.t().reshape(F_, self.C, self.N)

    def materialize(self):
        return self.Vt.reshape(self.C, self.N, *self.param_shape)


class LinearBiasFactor(DenseFactor):
    """Bias of a 2-d ``Linear`` layer: ``V_b^T = S``, so ``G_b = S S^T`` -- the very product the weight's Gram
    ``(Z Z^T) (.) (S S^T)`` is built from.  The reference computes it twice (``linear.py:72`` and, for the bias,
    ``base.py:124``); when both parameters sit in ONE group, ``G_W + G_b = (Z Z^T + 1) (.) (S S^T)`` and the
    weight factor's call takes the bias along (``fold_linear_bias`` sets the two flags)."""

    def __init__(self, S: Tensor, param_shape):
        super().__init__(S, param_shape)
        self.partner: Optional["LinearWeightFactor"] = None
        self.partner_param = None
        self.folded = False

    def gram_accum(self, G, fold=True):
        if not (fold and self.folded):
            super().gram_accum(G)


class LinearWeightFactor(Factor):
    """Structured factor of a 2-d ``Linear.weight``: ``V^T[(c,n), o, i] = S[c,n,o] Z[n,i]``
    (``linear.py:41-42``)."""

    def __init__(self, S: Tensor, Z: Tensor):
        self.C, self.N, self.n_out = S.shape
        self.n_in = Z.shape[1]
        self.param_shape = (self.n_out, self.n_in)
        self.S, self.Z = S, Z
        self.partner: Optional[LinearBiasFactor] = None
        self.partner_param = None
        self.fold_bias = False

    def _like(self):
        return self.S

    def gram_accum(self, G, fold=True, with_bias: Optional[bool] = None):
        if with_bias is None:
            with_bias = fold and self.fold_bias
        kernels.gram_linear_accum(G, self.S, self.Z, with_bias)

    def cross_accum(self, X, grad, with_bias: bool = False):
        if isinstance(grad, LinearWeightGrad):
            kernels.gram_cross_linear_accum(X, self.S, self.Z, grad.Dl, grad.Zg, with_bias)
        else:
            kernels.gram_cross_accum(X, self.materialize().reshape(self.R, -1), grad.dense())

    def backtransform(self, U, norm2):
        return kernels.backtransform_linear(U.reshape(U.shape[0], -1), self.S, self.Z, norm2)

    def v_apply(self, v):
        return kernels.v_apply_linear(v.reshape(-1), self.S, self.Z)

    def vt_mat_prod(self, M):
        return kernels.vt_mat_prod_linear(self.S, self.Z, M)

    def materialize(self):
        return kernels.v_emit_linear(self.S, self.Z)


def link_linear_factors(weight_factor: LinearWeightFactor, weight, bias_factor: LinearBiasFactor, bias) -> None:
    """Tell the two factors of one Linear layer about each other (and about each other's parameter).  Weak
    references: the two objects hold the layer's ``S`` and ``Z``, and a reference cycle would keep those alive until
    the cyclic garbage collector runs."""
    import weakref

    weight_factor.partner, weight_factor.partner_param = weakref.ref(bias_factor), bias
    bias_factor.partner, bias_factor.partner_param = weakref.ref(weight_factor), weight


def fold_linear_bias(factor: Factor, same_group) -> None:
    """Called by a Computation when a parameter's factor reaches its group: if the factor belongs to a Linear
    layer whose other parameter is in the same group (``same_group(param) -> bool``), the layer's two Gram
    contributions are assembled by ONE structured call (see ``LinearBiasFactor``)."""
    partner = getattr(factor, "partner", None)
    partner = partner() if partner is not None else None  # weak reference (link_linear_factors)
    if partner is None or factor.partner_param is None or not same_group(factor.partner_param):
        return
    weight, bias = (factor, partner) if isinstance(factor, LinearWeightFactor) else (partner, factor)
    weight.fold_bias = True
    bias.folded = True


class GradFactor:
    """Per-sample gradients of one parameter, ``[N_grad, *param.shape]``."""

    def dense(self) -> Tensor:  # [N_grad, D_p]
        raise NotImplementedError

    def materialize(self) -> Tensor:
        raise NotImplementedError


class DenseGrad(GradFactor):
    def __init__(self, g: Tensor, param_shape):
        self.g = g.reshape(g.shape[0], -1)
        self.param_shape = tuple(param_shape)

    def dense(self):
        return self.g

    def materialize(self):
        return self.g.reshape(self.g.shape[0], *self.param_shape)


class LinearWeightGrad(GradFactor):
    """``grad_batch[m] = Dl[m] (x) Zg[m]`` for a 2-d Linear weight ([BackPACK] BatchGrad)."""

    def __init__(self, Dl: Tensor, Zg: Tensor):
        self.Dl, self.Zg = Dl, Zg
        self.param_shape = (Dl.shape[1], Zg.shape[1])

    def dense(self):
        return self.materialize().reshape(self.Dl.shape[0], -1)

    def materialize(self):
        return kernels.v_emit_linear(self.Dl[None], self.Zg)[0]
