"""``Conv3d`` and ``ConvTranspose1d/2d/3d`` on the 2-d convolution kernels.

The reference reaches these layers through [BackPACK]'s ``ConvNDDerivatives`` / ``ConvTransposeNDDerivatives``
(``vivit/extensions/secondorder/vivit/__init__.py:84-101`` maps them to the generic ``param_mjp`` path of
``base.py:84-92``).  None of the BASELINE configs contains one, so they get no kernel of their own: a depth
index is a sum of 2-d problems, and a transposed convolution swaps the roles of the two operands of the 2-d
kernels.  Everything below is index plumbing around ``kernels.sqrt_backprop_conv2d`` (data gradient of a 2-d
convolution = a 2-d transposed convolution), ``kernels.v_emit_conv2d`` (per-sample weight Jacobian of a 2-d
convolution) and ``kernels.axpy_``; 1-d and 2-d transposed convolutions are 3-d ones of unit depth (and height).

Shapes: ``S [V, N, C_out, D', H', W']`` is the factor at the layer's OUTPUT, ``x [N, C_in, D, H, W]`` its input;
``stride`` / ``padding`` / ``dilation`` are 3-tuples ``(depth, height, width)``.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from vivit_b200 import kernels

Triple = Tuple[int, int, int]


def _accumulate(acc: Optional[Tensor], update: Tensor) -> Tensor:
    return update if acc is None else kernels.axpy_(acc, update)


def _stack(parts: List[Optional[Tensor]], dim: int, like_shape: Sequence[int], like: Tensor) -> Tensor:
    zeros = None
    out = []
    for p in parts:
        if p is None:  # no valid (output depth, kernel depth) combination reaches this slice
            if zeros is None:
                zeros = torch.zeros(*like_shape, dtype=like.dtype, device=like.device)
            p = zeros
        out.append(p)
    return torch.stack(out, dim=dim)


# ---------------------------------------------------------------------------------------------------------
# Conv3d: out[:, d'] = sum_a conv2d(x[:, d' * sd + a * dd - pd], W[:, :, a])
# ---------------------------------------------------------------------------------------------------------


def conv3d_backprop(S: Tensor, W: Tensor, in_dhw: Triple, stride: Triple, padding: Triple, dilation: Triple) -> Tensor:
    """``[V, N, Co, D', H', W'] -> [V, N, Ci, D, H, W]``: transposed Jacobian of a 3-d convolution with weight
    ``W [Co, Ci, kd, kh, kw]`` ([BackPACK] ``ConvNDDerivatives._jac_t_mat_prod``)."""
    V, N, _, Do = S.shape[:4]
    D, H, Wd = in_dhw
    kd = W.shape[2]
    slices: List[Optional[Tensor]] = [None] * D
    for a in range(kd):
        Wa = W[:, :, a].contiguous()
        for dp in range(Do):
            z = dp * stride[0] + a * dilation[0] - padding[0]
            if 0 <= z < D:
                part = kernels.sqrt_backprop_conv2d(S[:, :, :, dp].contiguous(), Wa, (H, Wd), stride[1:], padding[1:], dilation[1:])
                slices[z] = _accumulate(slices[z], part)
    return _stack(slices, 3, (V, N, W.shape[1], H, Wd), S)


def conv3d_emit(S: Tensor, x: Tensor, kernel: Triple, stride: Triple, padding: Triple, dilation: Triple) -> Tensor:
    """``V^T`` of a 3-d convolution weight, ``[V, N, Co, Ci, kd, kh, kw]`` ([BackPACK]
    ``ConvNDDerivatives.param_mjp`` with ``sum_batch=False``)."""
    V, N, Co, Do = S.shape[:4]
    Ci, D = x.shape[1], x.shape[2]
    taps: List[Optional[Tensor]] = []
    for a in range(kernel[0]):
        acc = None
        for dp in range(Do):
            z = dp * stride[0] + a * dilation[0] - padding[0]
            if 0 <= z < D:
                part = kernels.v_emit_conv2d(S[:, :, :, dp].contiguous(), x[:, :, z].contiguous(), kernel[1:], stride[1:], padding[1:], dilation[1:])
                acc = _accumulate(acc, part)
        taps.append(acc)
    return _stack(taps, 4, (V, N, Co, Ci, kernel[1], kernel[2]), S)


# ---------------------------------------------------------------------------------------------------------
# ConvTranspose: y[co, q] = sum_{ci, k, p: q = p s + k d - pad} x[ci, p] W[ci, co, k]
# ---------------------------------------------------------------------------------------------------------


def _conv2d_forward_of_factor(S2: Tensor, Wa: Tensor, out_hw, stride, padding, dilation) -> Tensor:
    """``out[ci, p] = sum_{co, k} Wa[ci, co, k] S2[co, p s + k d - pad]``: an ordinary 2-d convolution of the
    factor, which is the transposed Jacobian of a 2-d TRANSPOSED convolution.  The kernels hold the adjoint
    operation (the data gradient of a convolution), and at stride 1 the two coincide up to a flip of the filter:
    the stride-1 result is formed by the data-gradient kernel with the flipped, transposed filter and padding
    ``d (k - 1) - pad``, and the layer's stride picks every ``s``-th position of it."""
    kh, kw = Wa.shape[2:]
    flipped_pad = (dilation[0] * (kh - 1) - padding[0], dilation[1] * (kw - 1) - padding[1])
    if min(flipped_pad) < 0:
        raise NotImplementedError("transposed convolution with padding > dilation * (kernel_size - 1)")
    Hq, Wq = S2.shape[3:]
    full_hw = (Hq + 2 * padding[0] - dilation[0] * (kh - 1), Wq + 2 * padding[1] - dilation[1] * (kw - 1))
    Wc = Wa.transpose(0, 1).flip(2, 3).contiguous()  # [Co, Ci, kh, kw]: filter of the equivalent convolution
    full = kernels.sqrt_backprop_conv2d(S2, Wc, full_hw, (1, 1), flipped_pad, dilation)
    if stride[0] == 1 and stride[1] == 1:
        return full
    return full[..., :: stride[0], :: stride[1]][..., : out_hw[0], : out_hw[1]].contiguous()


def conv_transpose3d_backprop(
    S: Tensor, W: Tensor, in_dhw: Triple, stride: Triple, padding: Triple, dilation: Triple
) -> Tensor:
    """``[V, N, Co, D', H', W'] -> [V, N, Ci, D, H, W]`` for a transposed convolution with weight
    ``W [Ci, Co, kd, kh, kw]`` ([BackPACK] ``ConvTransposeNDDerivatives._jac_t_mat_prod``)."""
    V, N, _, Dq = S.shape[:4]
    D, H, Wd = in_dhw
    slices: List[Optional[Tensor]] = [None] * D
    for a in range(W.shape[2]):
        Wa = W[:, :, a]
        for pd in range(D):
            q = pd * stride[0] + a * dilation[0] - padding[0]
            if 0 <= q < Dq:
                part = _conv2d_forward_of_factor(S[:, :, :, q].contiguous(), Wa, (H, Wd), stride[1:], padding[1:], dilation[1:])
                slices[pd] = _accumulate(slices[pd], part)
    return _stack(slices, 3, (V, N, W.shape[0], H, Wd), S)


def conv_transpose3d_emit(S: Tensor, x: Tensor, kernel: Triple, stride: Triple, padding: Triple, dilation: Triple) -> Tensor:
    """``V^T`` of a transposed-convolution weight, ``[V, N, Ci, Co, kd, kh, kw]`` with
    ``V^T[v, n, ci, co, k] = sum_p x[n, ci, p] S[v, n, co, p s + k d - pad]``: the weight Jacobian of an ORDINARY
    convolution whose input is the factor row ``S[v, n]`` and whose output-side operand is ``x[n]`` -- the emit
    kernel with the operands swapped and every factor row as its own "sample"."""
    V, N, Co, Dq = S.shape[:4]
    Ci, D = x.shape[1], x.shape[2]
    rows = V * N
    taps: List[Optional[Tensor]] = []
    for a in range(kernel[0]):
        acc = None
        for pd in range(D):
            q = pd * stride[0] + a * dilation[0] - padding[0]
            if 0 <= q < Dq:
                x_rows = x[:, :, pd].unsqueeze(0).expand(V, *x[:, :, pd].shape).reshape(1, rows, Ci, *x.shape[3:])
                s_rows = S[:, :, :, q].reshape(rows, Co, *S.shape[4:])
                part = kernels.v_emit_conv2d(x_rows, s_rows, kernel[1:], stride[1:], padding[1:], dilation[1:])
                acc = _accumulate(acc, part)
        taps.append(None if acc is None else acc.reshape(V, N, Ci, Co, kernel[1], kernel[2]))
    return _stack(taps, 4, (V, N, Ci, Co, kernel[1], kernel[2]), S)


def lift_to_3d(t: Tensor, nd: int) -> Tensor:
    """``[..., *spatial(nd)] -> [..., D, H, W]`` with unit depth (and height) for 2-d (1-d) layers."""
    for _ in range(3 - nd):
        t = t.unsqueeze(-nd - 1)
    return t


def triple(values, nd: int, fill: int) -> Triple:
    values = (values,) * nd if isinstance(values, int) else tuple(values)
    return (fill,) * (3 - nd) + values
