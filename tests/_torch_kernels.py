"""Plain-torch test double of ``vivit_b200.kernels`` (same signatures, any device).

TEST INFRASTRUCTURE ONLY.  It lives under ``tests/`` and is never importable
from the ``vivit_b200`` package: the product has no CPU/PyTorch fallback.  Two uses:

* ``-m "not gpu"`` tests install it with ``monkeypatch`` so the host-side logic
  (hook engine, parameter groups, sharding, error contract) can be exercised on
  a machine without a GPU;
* ``-m gpu`` tests call it on CUDA tensors in float64 as the per-kernel
  reference for each entry point of the C ABI.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import einsum

(ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_DROPOUT, ACT_MUL, ACT_LEAKY_RELU, ACT_ELU, ACT_SELU,
 ACT_LOGSIGMOID) = range(9)

NAMES = [
    "loss_sqrt_hessian_ce", "loss_sqrt_hessian_ce_mc", "loss_sqrt_hessian_mse", "scale_", "axpy_",
    "sqrt_backprop_linear", "sqrt_backprop_conv2d", "sqrt_backprop_elementwise",
    "sqrt_backprop_maxpool2d", "sqrt_backprop_avgpool2d", "v_emit_conv2d", "v_emit_bias",
    "v_emit_linear", "gemm", "gram_dense_accum", "gram_cross_accum", "gram_linear_accum",
    "gram_cross_linear_accum", "syevj", "syevj_batched", "filter_nonzero", "backtransform_dense",
    "backtransform_linear", "vt_mat_prod_linear", "scale_rows_rsqrt", "dirderiv_epilogue",
    "newton_coeff", "v_apply_dense", "v_apply_linear", "launch_count", "center_rows", "maxpool2d_argmax",
]



def launch_count():
    return 0


def install(monkeypatch):
    """Route ``vivit_b200.kernels`` through this module (tests only)."""
    import sys

    import vivit_b200.kernels as k

    me = sys.modules[__name__]
    for name in NAMES:
        monkeypatch.setattr(k, name, getattr(me, name))


def _sub(t, sub):
    return t if sub is None else t[sub]


def loss_sqrt_hessian_ce(logits, sub, mean):
    n_total, C = logits.shape
    tau = _sub(F.softmax(logits, 1), sub).sqrt()
    eye = torch.eye(C, dtype=logits.dtype, device=logits.device)
    S = einsum("nc,vnc->vnc", tau, eye[:, None, :] - einsum("nv,nc->vnc", tau, tau))
    return S / math.sqrt(n_total) if mean else S


def loss_sqrt_hessian_ce_mc(logits, sub, class_ids, mean):
    n_total, C = logits.shape
    p = _sub(F.softmax(logits, 1), sub)
    M = class_ids.shape[0]
    S = (p[None] - F.one_hot(class_ids, C).to(p.dtype)) / math.sqrt(M)
    return S / math.sqrt(n_total) if mean else S


def loss_sqrt_hessian_mse(n_sub, C, scale, like):
    eye = torch.eye(C, dtype=like.dtype, device=like.device) * scale
    return eye[:, None, :].expand(C, n_sub, C).contiguous()


def scale_(t, alpha):
    return t.mul_(alpha)


def axpy_(y, x, alpha=1.0):
    return y.add_(x, alpha=alpha)


def center_rows(g, inplace=False):
    mean = g.mean(0)
    return g.sub_(mean) if inplace else g - mean


def sqrt_backprop_linear(S, W):
    return S @ W


def sqrt_backprop_conv2d(S, W, in_hw, stride, padding, dilation):
    V, N = S.shape[:2]
    x = torch.zeros(V * N, W.shape[1], *in_hw, dtype=S.dtype, device=S.device, requires_grad=True)
    with torch.enable_grad():
        y = F.conv2d(x, W, None, stride, padding, dilation)
        (g,) = torch.autograd.grad(y, x, S.reshape(V * N, *S.shape[2:]))
    return g.reshape(V, N, *g.shape[1:])


def sqrt_backprop_elementwise(S, ref, act, scale=1.0):
    if act == ACT_RELU:
        d = (ref > 0).to(S.dtype)
    elif act == ACT_SIGMOID:
        d = ref * (1 - ref)
    elif act == ACT_TANH:
        d = 1 - ref**2
    elif act == ACT_DROPOUT:
        d = (ref != 0).to(S.dtype) * scale
    elif act == ACT_LEAKY_RELU:
        d = torch.where(ref > 0, torch.ones_like(ref), torch.full_like(ref, scale))
    elif act == ACT_ELU:
        d = torch.where(ref > 0, torch.ones_like(ref), scale * ref.exp())
    elif act == ACT_SELU:
        d = 1.0507009873554804934193349852946 * torch.where(
            ref > 0, torch.ones_like(ref), 1.6732632423543772848170429916717 * ref.exp())
    elif act == ACT_LOGSIGMOID:
        d = 1.0 / (1.0 + ref.exp())
    else:
        d = ref
    return (S.reshape(-1, ref.numel()) * d.reshape(1, -1)).reshape(S.shape)


def maxpool2d_argmax(x, kernel, stride, padding, dilation, ceil_mode=False):
    return F.max_pool2d(x, kernel, stride, padding, dilation, ceil_mode, return_indices=True)[1]


def sqrt_backprop_maxpool2d(S, argmax, in_hw, kernel, stride, padding, dilation):
    V, N, ch = S.shape[:3]
    h, w = in_hw
    out = torch.zeros(V, N, ch, h * w, dtype=S.dtype, device=S.device)
    out.scatter_add_(3, argmax.reshape(1, N, ch, -1).expand(V, -1, -1, -1), S.reshape(V, N, ch, -1))
    return out.reshape(V, N, ch, h, w)


def sqrt_backprop_avgpool2d(S, in_hw, kernel, stride, padding):
    V, N, ch = S.shape[:3]
    x = torch.zeros(V * N, ch, *in_hw, dtype=S.dtype, device=S.device, requires_grad=True)
    with torch.enable_grad():
        y = F.avg_pool2d(x, kernel, stride, padding)
        (g,) = torch.autograd.grad(y, x, S.reshape(V * N, *S.shape[2:]))
    return g.reshape(V, N, ch, *in_hw)


def v_emit_conv2d(S, X, kernel, stride, padding, dilation):
    V, N, co = S.shape[:3]
    if co == 0:
        return S.new_zeros(V, N, 0, X.shape[1], *kernel)
    cols = F.unfold(X, kernel, dilation=dilation, padding=padding, stride=stride)
    vt = einsum("vnox,njx->vnoj", S.reshape(V, N, co, -1), cols)
    return vt.reshape(V, N, co, X.shape[1], *kernel)


def v_emit_bias(S):
    return S.flatten(3).sum(3) if S.dim() > 3 else S.clone()


def v_emit_linear(S, Z):
    return einsum("vno,ni->vnoi", S, Z)


def gemm(A, B, trans_a=False, trans_b=False, out=None, alpha=1.0, beta=0.0):
    a = A.transpose(-1, -2) if trans_a else A
    b = B if trans_b else B.transpose(-1, -2)
    res = alpha * (a @ b)
    if out is None:
        return res
    out.mul_(beta).add_(res)
    return out


def gram_dense_accum(G, V):
    G.add_(V @ V.t())
    return G


def gram_cross_accum(X, V, g):
    X.add_(V @ g.t())
    return X


def gram_linear_accum(G, S, Z, with_bias):
    C, N, _ = S.shape
    P = Z @ Z.t() + (1.0 if with_bias else 0.0)
    s2 = einsum("cno,dmo->cndm", S, S)
    G.add_((s2 * P[None, :, None, :]).reshape(C * N, C * N))
    return G


def gram_cross_linear_accum(X, S, Z, Dl, Zg, with_bias):
    C, N, _ = S.shape
    P = Z @ Zg.t() + (1.0 if with_bias else 0.0)
    X.add_((einsum("cno,mo->cnm", S, Dl) * P[None]).reshape(C * N, -1))
    return X


class SyevjNotConverged(RuntimeError):
    pass


def syevj(G, vectors=True, return_info=False):
    sym = torch.triu(G) + torch.triu(G, 1).t()
    info = ({"sweeps": 0, "converged": True},) if return_info else ()
    if vectors:
        return tuple(torch.linalg.eigh(sym)) + info
    return (torch.linalg.eigvalsh(sym), None) + info


def syevj_batched(G, vectors=True, return_info=False):
    sym = torch.triu(G) + torch.triu(G, 1).transpose(-1, -2)
    infos = ([{"sweeps": 0, "converged": True} for _ in range(G.shape[0])],) if return_info else ()
    if vectors:
        return tuple(torch.linalg.eigh(sym)) + infos
    return (torch.linalg.eigvalsh(sym), None) + infos


def filter_nonzero(evals, atol=1e-7, rtol=1e-5):
    return torch.isclose(evals, torch.zeros_like(evals), rtol=rtol, atol=atol).logical_not()


def backtransform_dense(U, V, norm2):
    E = U @ V
    if norm2 is not None:
        norm2.add_((E.double() ** 2).sum(1))
    return E


def backtransform_linear(U, S, Z, norm2):
    C, N, _ = S.shape
    E = einsum("cno,kcn,ni->koi", S, U.reshape(-1, C, N), Z)
    if norm2 is not None:
        norm2.add_((E.double() ** 2).flatten(1).sum(1))
    return E


def vt_mat_prod_linear(S, Z, M):
    return einsum("cno,foi,ni->fcn", S, M, Z)


def scale_rows_rsqrt(E, norm2):
    shape = (-1,) + (1,) * (E.dim() - 1)
    E.mul_((1.0 / norm2.sqrt()).to(E.dtype).reshape(shape))
    return E


def dirderiv_epilogue(G, X, U, evals, C, N_ggn, N):
    corr2 = N / N_ggn
    corr = math.sqrt(corr2)
    gammas = (corr * N) * (X.t() @ U) / evals.sqrt()
    W = (corr2 * G) @ U
    lambdas = N_ggn * (W.reshape(C, N_ggn, -1) ** 2).sum(0) / evals
    return gammas, lambdas


def newton_coeff(U, gammas, lambdas, deltas, evals, corr):
    coef = -gammas.mean(0) / (lambdas.mean(0) + deltas) / evals.sqrt()
    return (U @ coef) * corr


def v_apply_dense(v, V):
    return v @ V


def v_apply_linear(v, S, Z):
    C, N, _ = S.shape
    return einsum("cno,cn,ni->oi", S, v.reshape(C, N), Z)
