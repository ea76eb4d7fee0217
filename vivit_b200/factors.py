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


_CONV_STREAM_BYTES: Optional[int] = None


def set_conv_factor_streaming(chunk_bytes: Optional[int]) -> None:
    """``chunk_bytes`` (e.g. ``48 << 20``): conv weights get a ``StreamedConvFactor`` -- ``V_p^T`` is emitted in
    output-channel chunks of at most that many bytes whenever it is needed and never exists as a whole.  ``None``
    (default): the factor is materialised once (``DenseFactor``), which is faster when it is used more than once and
    fits.  ``VVT_STREAM_CONV_FACTOR_MB`` in the environment sets the default."""
    global _CONV_STREAM_BYTES
    _CONV_STREAM_BYTES = None if chunk_bytes is None else max(1, int(chunk_bytes))


def conv_factor_streaming() -> Optional[int]:
    if _CONV_STREAM_BYTES is not None:
        return _CONV_STREAM_BYTES
    import os

    mb = os.environ.get("VVT_STREAM_CONV_FACTOR_MB")
    return int(float(mb) * (1 << 20)) if mb else None


class StreamedConvFactor(Factor):
    """``V_p^T`` of a convolution weight that is never materialised as a whole (north star (2), VERDICT n2 at the
    host level): the factor ``S [C, N, Co, Ho, Wo]`` of the layer output and the layer input ``x`` are kept, and every
    use emits ``V_p^T`` for a chunk of output channels (``vvt_v_emit_conv2d`` on a channel slice, the call the
    parameter-sharded path makes per rank), consumes it (Gram / cross term / back-transform) and drops it.  A chunk
    of a few tens of MB stays in the 126 MB L2 between the emit and its consumer.  Costs one emit per use instead of
    one per pass -- DESIGN section 4 has the arithmetic of why the DEFAULT keeps the factor materialised; this is for
    factors that should not occupy ``R x D_p`` of HBM (``reference: base.py:84-92`` materialises them always)."""

    def __init__(self, S: Tensor, x: Tensor, kernel, geom, param_shape, chunk_bytes: int):
        self.C, self.N = S.shape[:2]
        self.param_shape = tuple(param_shape)  # (Co_own, Ci, kh, kw); (Co_own, Ci, k) for a 1-d convolution
        self.S, self.x, self.kernel, self.geom = S, x, tuple(kernel), tuple(geom)
        self.per_channel = 1
        for d in self.param_shape[1:]:
            self.per_channel *= d
        row_bytes = self.C * self.N * self.per_channel * S.element_size()
        self.step = max(1, int(chunk_bytes) // max(1, row_bytes))

    def _like(self):
        return self.S

    def chunks(self):
        """``(lo, hi, Vt [R, (hi - lo) * Ci * kh * kw])`` over the output channels."""
        co = self.param_shape[0]
        for lo in range(0, co, self.step):
            hi = min(co, lo + self.step)
            S = self.S if (lo, hi) == (0, co) else self.S[:, :, lo:hi].contiguous()
            Vt = kernels.v_emit_conv2d(S, self.x, self.kernel, *self.geom)
            yield lo, hi, Vt.reshape(self.R, -1)

    def _cols(self, lo, hi):
        return slice(lo * self.per_channel, hi * self.per_channel)

    def gram_accum(self, G, fold=True):
        for _, _, Vt in self.chunks():
            kernels.gram_dense_accum(G, Vt)

    def cross_accum(self, X, grad):
        g = grad.dense()
        for lo, hi, Vt in self.chunks():
            kernels.gram_cross_accum(X, Vt, g[:, self._cols(lo, hi)].contiguous())

    def backtransform(self, U, norm2):
        U = U.reshape(U.shape[0], -1)
        parts = [kernels.backtransform_dense(U, Vt, norm2) for _, _, Vt in self.chunks()]  # norm2 accumulates
        return torch.cat(parts, dim=1).reshape(U.shape[0], *self.param_shape)

    def v_apply(self, v):
        parts = [kernels.v_apply_dense(v.reshape(-1), Vt) for _, _, Vt in self.chunks()]
        return torch.cat(parts).reshape(self.param_shape)

    def vt_mat_prod(self, M):
        F_ = M.shape[0]
        M = M.reshape(F_, -1)
        out = torch.zeros(self.R, F_, dtype=self.S.dtype, device=self.S.device)
        for lo, hi, Vt in self.chunks():
            kernels.gram_cross_accum(out, Vt, M[:, self._cols(lo, hi)].contiguous())
        return out.t().reshape(F_, self.C, self.N)

    def materialize(self):
        return torch.cat([Vt for _, _, Vt in self.chunks()], dim=1).reshape(self.C, self.N, *self.param_shape)


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
