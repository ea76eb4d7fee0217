"""Python entry points of the sm_100a kernels (thin wrappers over the C ABI).

Every function takes CUDA tensors, allocates its outputs with torch (device
memory and streams are torch's job), and launches hand-written kernels from
``libvivit_b200.so`` on torch's current stream.  There is no fallback: a CPU
tensor or a missing library raises.

Shapes use the reference's symbols: ``C`` classes (or MC samples), ``N``
samples of the (sub-sampled) batch, ``R = C*N`` Gram dimension, ``K`` kept
directions.
"""

from __future__ import annotations

import contextlib
import ctypes
import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from vivit_b200 import _lib

(ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_DROPOUT, ACT_MUL, ACT_LEAKY_RELU, ACT_ELU, ACT_SELU,
 ACT_LOGSIGMOID) = range(9)

_DTYPES = {torch.float32: 0, torch.float64: 1}


def _dt(t: Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"vivit_b200 kernels support float32/float64, got {t.dtype}") from None


def _chk(*tensors: Optional[Tensor]) -> None:
    """Operands of one kernel call: CUDA, contiguous, one device, ONE floating dtype (the C ABI takes a
    single dtype enum per call: a float32 factor against a float64 weight would be reinterpreted bitwise),
    integer operands int64 (index arrays) or uint8 (masks)."""
    dev = None
    fdt = None
    for t in tensors:
        if t is None:
            continue
        if t.is_floating_point():
            if fdt is None:
                fdt = t.dtype
            elif t.dtype != fdt:
                raise TypeError(f"kernel operands have different floating dtypes: {fdt} and {t.dtype}")
        elif t.dtype not in (torch.int64, torch.uint8, torch.bool):
            raise TypeError(f"integer kernel operands must be int64 (indices), got {t.dtype}")
        if not t.is_cuda:
            raise _lib.KernelLibraryError(
                "vivit_b200 kernels run on CUDA tensors only (there is no CPU fallback)"
            )
        if not t.is_contiguous():
            raise ValueError("kernel operands must be contiguous")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("kernel operands live on different devices")


def _p(t: Optional[Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_NO_SWITCH = contextlib.nullcontext()


def _stream_handle(device: torch.device) -> int:
    """``cudaStream_t`` of torch's current stream on ``device``.  (The raw query is ~20x cheaper than building a
    ``torch.cuda.Stream`` object; a curvature step makes a few hundred kernel calls and the GPU idles at the start
    of a step while the host catches up.)"""
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


def _stream(t: Tensor):
    return ctypes.c_void_p(_stream_handle(t.device))


def _on(device: torch.device):
    """Context that makes ``device`` current for a kernel call -- nothing to do when it already is."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)


# Workspace arena: one byte buffer per (device, stream), grown on demand and reused by every kernel entry point that
# needs scratch memory.  Calls on one stream execute in order, so a later call may overwrite the scratch of an earlier
# one; nothing a caller receives ever aliases it.  (Fresh ``torch.empty`` workspaces of 50 MB ... 3 GB per call made
# the caching allocator split and re-allocate blocks: a step of config 3 took 16 or 41 ms depending on the
# allocator's history.)  ``release_workspaces()`` hands the memory back.
_WORKSPACES = {}


def _ws(nbytes: int, like: Tensor) -> Tensor:
    nbytes = max(int(nbytes), 16)
    key = (like.device.index, _stream_handle(like.device))
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        _WORKSPACES.pop(key, None)
        buf = None  # free the old block before asking for the larger one
        buf = torch.empty(nbytes + nbytes // 8, dtype=torch.uint8, device=like.device)
        _WORKSPACES[key] = buf
    return buf[:nbytes]


def release_workspaces() -> None:
    """Drop the cached scratch buffers (they are re-created on demand)."""
    _WORKSPACES.clear()


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def launch_count() -> int:
    """Kernels launched by the library so far (host-side counter)."""
    return int(_lib.load().vvt_launch_count())


# --------------------------------------------------------------------------
# (1) loss-Hessian factor
# --------------------------------------------------------------------------


def loss_sqrt_hessian_ce(logits: Tensor, sub: Optional[Tensor], mean: bool) -> Tensor:
    """``[C, N_sub, C]`` exact factor of the softmax cross-entropy Hessian."""
    logits = _c(logits)
    _chk(logits, sub)
    n_total, C = logits.shape
    n_sub = n_total if sub is None else sub.numel()
    S = torch.empty(C, n_sub, C, dtype=logits.dtype, device=logits.device)
    scale = 1.0 / math.sqrt(n_total) if mean else 1.0
    with _on(logits.device):
        st = _lib.load().vvt_loss_sqrt_hessian_ce(
            _p(S), _p(logits), _p(sub), n_total, n_sub, C, scale, _dt(logits), _stream(logits)
        )
    _lib.check(st, "vvt_loss_sqrt_hessian_ce")
    return S


def loss_sqrt_hessian_ce_mc(
    logits: Tensor, sub: Optional[Tensor], class_ids: Tensor, mean: bool
) -> Tensor:
    """``[M, N_sub, C]`` sampled factor for class ids ``[M, N_sub]``."""
    logits, class_ids = _c(logits), _c(class_ids)
    _chk(logits, sub, class_ids)
    n_total, C = logits.shape
    M, n_sub = class_ids.shape
    S = torch.empty(M, n_sub, C, dtype=logits.dtype, device=logits.device)
    scale = 1.0 / math.sqrt(M) / (math.sqrt(n_total) if mean else 1.0)
    with _on(logits.device):
        st = _lib.load().vvt_loss_sqrt_hessian_ce_mc(
            _p(S), _p(logits), _p(sub), _p(class_ids), n_total, n_sub, C, M, scale,
            _dt(logits), _stream(logits),
        )
    _lib.check(st, "vvt_loss_sqrt_hessian_ce_mc")
    return S


def loss_sqrt_hessian_mse(n_sub: int, C: int, scale: float, like: Tensor) -> Tensor:
    """``[C, N_sub, C]`` with ``S[v,n,c] = scale * delta_vc``."""
    _chk(like)
    S = torch.empty(C, n_sub, C, dtype=like.dtype, device=like.device)
    with _on(like.device):
        st = _lib.load().vvt_loss_sqrt_hessian_mse(_p(S), n_sub, C, scale, _dt(like), _stream(like))
    _lib.check(st, "vvt_loss_sqrt_hessian_mse")
    return S


def scale_(t: Tensor, alpha: float) -> Tensor:
    _chk(t)
    with _on(t.device):
        st = _lib.load().vvt_scale(_p(t), t.numel(), float(alpha), _dt(t), _stream(t))
    _lib.check(st, "vvt_scale")
    return t


def axpy_(y: Tensor, x: Tensor, alpha: float = 1.0) -> Tensor:
    """``y += alpha * x`` in place (both contiguous, same shape and dtype)."""
    _chk(y, x)
    if y.shape != x.shape or not (y.is_contiguous() and x.is_contiguous()):
        raise ValueError("axpy_ needs contiguous tensors of one shape")
    with _on(y.device):
        st = _lib.load().vvt_axpy(_p(y), _p(x), y.numel(), float(alpha), _dt(y), _stream(y))
    _lib.check(st, "vvt_axpy")
    return y


def nccl_allreduce_gram(comm_ptr: int, G: Tensor, X: Optional[Tensor] = None, alpha: float = 1.0) -> None:
    """``G`` (and ``X``) ``<- alpha * sum over ranks``, in place, one NCCL call group on the current stream;
    ``comm_ptr`` is the raw ``ncclComm_t`` of the process group (``ProcessGroupNCCL._comm_ptr()``)."""
    _chk(G, X)
    with _on(G.device):
        st = _lib.load().vvt_nccl_allreduce_gram(
            ctypes.c_void_p(comm_ptr), _p(G), G.numel(), _p(X), 0 if X is None else X.numel(), float(alpha),
            _dt(G), _stream(G),
        )
    _lib.check(st, "vvt_nccl_allreduce_gram")


def center_rows(g: Tensor, inplace: bool = False) -> Tensor:
    """``g [N, ...] - g.mean(0)`` (per-sample gradients minus their mean); ``inplace`` overwrites ``g``."""
    if inplace and not g.is_contiguous():
        raise ValueError("in-place centering needs a contiguous tensor")
    g = _c(g)
    _chk(g)
    out = g if inplace else torch.empty_like(g)
    N = g.shape[0]
    D = g.numel() // N if N else 0
    with _on(g.device):
        st = _lib.load().vvt_center_rows(_p(out), _p(g), N, D, _dt(g), _stream(g))
    _lib.check(st, "vvt_center_rows")
    return out


# --------------------------------------------------------------------------
# (1) factor back-propagation
# --------------------------------------------------------------------------


def sqrt_backprop_linear(S: Tensor, W: Tensor) -> Tensor:
    """``[..., out] x [out, in] -> [..., in]``."""
    S, W = _c(S), _c(W)
    _chk(S, W)
    n_out, n_in = W.shape
    rows = S.numel() // n_out
    out = torch.empty(*S.shape[:-1], n_in, dtype=S.dtype, device=S.device)
    with _on(S.device):
        st = _lib.load().vvt_sqrt_backprop_linear(
            _p(out), _p(S), _p(W), rows, n_out, n_in, _dt(S), _stream(S)
        )
    _lib.check(st, "vvt_sqrt_backprop_linear")
    return out


def sqrt_backprop_conv2d(
    S: Tensor, W: Tensor, in_hw: Tuple[int, int], stride, padding, dilation
) -> Tensor:
    """``[V, N, Co, Ho, Wo] -> [V, N, Ci, H, W]`` (data gradient of the convolution)."""
    S, W = _c(S), _c(W)
    _chk(S, W)
    V, N, co, ho, wo = S.shape
    _, ci, kh, kw = W.shape
    h, w = in_hw
    out = torch.empty(V, N, ci, h, w, dtype=S.dtype, device=S.device)
    lib = _lib.load()
    ws = _ws(lib.vvt_conv2d_workspace_bytes(1, V, N, co, ho, wo, ci, kh, kw, _dt(S)), S)
    with _on(S.device):
        st = lib.vvt_sqrt_backprop_conv2d(
            _p(out), _p(S), _p(W), V * N, co, ho, wo, ci, h, w, kh, kw,
            stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1],
            _p(ws), ws.numel(), _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_sqrt_backprop_conv2d")
    return out


def sqrt_backprop_elementwise(S: Tensor, ref: Tensor, act: int, scale: float = 1.0) -> Tensor:
    """``S[v, n, ...] * J(ref[n, ...])`` for an element-wise layer."""
    S, ref = _c(S), _c(ref)
    _chk(S, ref)
    n_feat = ref.numel()
    V = S.numel() // n_feat
    out = torch.empty_like(S)
    with _on(S.device):
        st = _lib.load().vvt_sqrt_backprop_elementwise(
            _p(out), _p(S), _p(ref), V, n_feat, act, float(scale), _dt(S), _stream(S)
        )
    _lib.check(st, "vvt_sqrt_backprop_elementwise")
    return out


def pool_output_size(size: int, kernel: int, stride: int, padding: int, dilation: int, ceil_mode: bool) -> int:
    """Output extent of a pooling window sweep (torch's rule; in ceil mode the last window must start inside the
    input or its left padding)."""
    span = size + 2 * padding - dilation * (kernel - 1) - 1 + (stride - 1 if ceil_mode else 0)
    out = span // stride + 1
    if ceil_mode and (out - 1) * stride >= size + padding:
        out -= 1
    return out


def maxpool2d_argmax(x: Tensor, kernel, stride, padding, dilation, ceil_mode: bool = False) -> Tensor:
    """Arg-max positions ``[N, ch, Ho, Wo]`` (``int64``, flat into ``H*W``) of ``max_pool2d(x [N, ch, H, W])``,
    chosen as torch chooses them (first of equal values).  The index plumbing of the max-pool Jacobian."""
    x = _c(x)
    _chk(x)
    N, ch, h, w = x.shape
    ho = pool_output_size(h, kernel[0], stride[0], padding[0], dilation[0], ceil_mode)
    wo = pool_output_size(w, kernel[1], stride[1], padding[1], dilation[1], ceil_mode)
    if ho <= 0 or wo <= 0:
        raise ValueError(f"max-pool window {tuple(kernel)} does not fit the input {(h, w)}")
    idx = torch.empty(N, ch, ho, wo, dtype=torch.int64, device=x.device)
    with _on(x.device):
        st = _lib.load().vvt_maxpool2d_argmax(
            _p(idx), _p(x), N * ch, ho, wo, h, w, kernel[0], kernel[1], stride[0], stride[1],
            padding[0], padding[1], dilation[0], dilation[1], _dt(x), _stream(x),
        )
    _lib.check(st, "vvt_maxpool2d_argmax")
    return idx


def sqrt_backprop_maxpool2d(
    S: Tensor, argmax: Tensor, in_hw, kernel, stride, padding, dilation
) -> Tensor:
    """``[V, N, ch, Ho, Wo] -> [V, N, ch, H, W]``; ``argmax`` from ``max_pool2d(return_indices=True)``."""
    S, argmax = _c(S), _c(argmax)
    _chk(S, argmax)
    V, N, ch, ho, wo = S.shape
    h, w = in_hw
    out = torch.empty(V, N, ch, h, w, dtype=S.dtype, device=S.device)
    with _on(S.device):
        st = _lib.load().vvt_sqrt_backprop_maxpool2d(
            _p(out), _p(S), _p(argmax), V, N, ch, ho, wo, h, w, kernel[0], kernel[1],
            stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1],
            _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_sqrt_backprop_maxpool2d")
    return out


def sqrt_backprop_avgpool2d(S: Tensor, in_hw, kernel, stride, padding) -> Tensor:
    S = _c(S)
    _chk(S)
    V, N, ch, ho, wo = S.shape
    h, w = in_hw
    out = torch.empty(V, N, ch, h, w, dtype=S.dtype, device=S.device)
    with _on(S.device):
        st = _lib.load().vvt_sqrt_backprop_avgpool2d(
            _p(out), _p(S), V * N, ch, ho, wo, h, w, kernel[0], kernel[1],
            stride[0], stride[1], padding[0], padding[1], _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_sqrt_backprop_avgpool2d")
    return out


# --------------------------------------------------------------------------
# (1) emitting V^T in [R, D_p] layout
# --------------------------------------------------------------------------


def v_emit_conv2d(S: Tensor, X: Tensor, kernel, stride, padding, dilation) -> Tensor:
    """``S [V,N,Co,Ho,Wo]``, ``X [N,Ci,H,W]`` -> ``[V, N, Co, Ci, kh, kw]``."""
    S, X = _c(S), _c(X)
    _chk(S, X)
    V, N, co, ho, wo = S.shape
    _, ci, h, w = X.shape
    kh, kw = kernel
    Vt = torch.empty(V, N, co, ci, kh, kw, dtype=S.dtype, device=S.device)
    if Vt.numel() == 0:  # an empty shard of the output channels (more ranks than channels)
        return Vt
    lib = _lib.load()
    ws = _ws(lib.vvt_conv2d_workspace_bytes(0, V, N, co, ho, wo, ci, kh, kw, _dt(S)), S)
    with _on(S.device):
        st = lib.vvt_v_emit_conv2d(
            _p(Vt), _p(S), _p(X), V, N, co, ho, wo, ci, h, w, kh, kw,
            stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1],
            _p(ws), ws.numel(), _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_v_emit_conv2d")
    return Vt


def v_emit_bias(S: Tensor) -> Tensor:
    """``[V, N, Co, *spatial] -> [V, N, Co]`` (sum over spatial positions)."""
    S = _c(S)
    _chk(S)
    V, N, co = S.shape[:3]
    Vt = torch.empty(V, N, co, dtype=S.dtype, device=S.device)
    if Vt.numel() == 0:  # an empty shard of the output channels
        return Vt
    spatial = S.numel() // (V * N * co)
    with _on(S.device):
        st = _lib.load().vvt_v_emit_bias(_p(Vt), _p(S), V * N, co, spatial, _dt(S), _stream(S))
    _lib.check(st, "vvt_v_emit_bias")
    return Vt


def v_emit_linear(S: Tensor, Z: Tensor) -> Tensor:
    """Materialise ``[V, N, out, in]`` from ``S [V,N,out]`` and ``Z [N,in]``."""
    S, Z = _c(S), _c(Z)
    _chk(S, Z)
    V, N, n_out = S.shape
    n_in = Z.shape[1]
    Vt = torch.empty(V, N, n_out, n_in, dtype=S.dtype, device=S.device)
    with _on(S.device):
        st = _lib.load().vvt_v_emit_linear(
            _p(Vt), _p(S), _p(Z), V, N, n_out, n_in, _dt(S), _stream(S)
        )
    _lib.check(st, "vvt_v_emit_linear")
    return Vt


# --------------------------------------------------------------------------
# (2) Gram assembly
# --------------------------------------------------------------------------


def gemm(
    A: Tensor, B: Tensor, trans_a: bool = False, trans_b: bool = False,
    out: Optional[Tensor] = None, alpha: float = 1.0, beta: float = 0.0,
) -> Tensor:
    """``out = alpha * op(A) @ op(B)^T + beta * out`` with ``op(B)`` given as ``[N, K]``
    (``trans_b=False``) or ``[K, N]`` (``trans_b=True``); 2-d or batched 3-d operands."""
    A, B = _c(A), _c(B)
    batched = A.dim() == 3
    a2 = A.shape[-2:]
    b2 = B.shape[-2:]
    M, K = (a2[1], a2[0]) if trans_a else (a2[0], a2[1])
    N = b2[1] if trans_b else b2[0]
    batch = A.shape[0] if batched else 1
    if out is None:
        shape = (batch, M, N) if batched else (M, N)
        out = torch.empty(shape, dtype=A.dtype, device=A.device)
        beta = 0.0
    _chk(A, B, out)
    lib = _lib.load()
    ws = _ws(lib.vvt_gram_workspace_bytes(M, N, K, _dt(A)) if batch == 1 else 0, A)
    with _on(A.device):
        st = lib.vvt_gemm(
            _p(out), _p(A), _p(B), M, N, K, int(trans_a), int(trans_b),
            a2[1], b2[1], N, float(alpha), float(beta), batch,
            a2[0] * a2[1], b2[0] * b2[1] if B.dim() == 3 else 0, M * N,
            _p(ws), ws.numel(), _dt(A), _stream(A),
        )
    _lib.check(st, "vvt_gemm")
    return out


def gram_dense_accum(G: Tensor, V: Tensor) -> Tensor:
    """``G [R,R] += V V^T`` for ``V [R, D]``."""
    V = _c(V)
    _chk(G, V)
    R, D = V.shape
    lib = _lib.load()
    ws = _ws(lib.vvt_gram_workspace_bytes(R, R, D, _dt(V)), V)
    with _on(V.device):
        st = lib.vvt_gram_dense_accum(_p(G), _p(V), R, D, _p(ws), ws.numel(), _dt(V), _stream(V))
    _lib.check(st, "vvt_gram_dense_accum")
    return G


def gram_cross_accum(X: Tensor, V: Tensor, g: Tensor) -> Tensor:
    """``X [R, n_g] += V g^T`` for ``V [R, D]``, ``g [n_g, D]``."""
    V, g = _c(V), _c(g)
    _chk(X, V, g)
    R, D = V.shape
    n_g = g.shape[0]
    lib = _lib.load()
    ws = _ws(lib.vvt_gram_workspace_bytes(R, n_g, D, _dt(V)), V)
    with _on(V.device):
        st = lib.vvt_gram_cross_accum(
            _p(X), _p(V), _p(g), R, n_g, D, _p(ws), ws.numel(), _dt(V), _stream(V)
        )
    _lib.check(st, "vvt_gram_cross_accum")
    return X


def gram_linear_accum(G: Tensor, S: Tensor, Z: Tensor, with_bias: bool) -> Tensor:
    """``G += (Z Z^T + with_bias) (.) (S S^T)``; ``S [C, N, out]``, ``Z [N, in]``."""
    S, Z = _c(S), _c(Z)
    _chk(G, S, Z)
    C, N, n_out = S.shape
    n_in = Z.shape[1]
    lib = _lib.load()
    ws = _ws(lib.vvt_gram_linear_workspace_bytes(C, N, n_out, n_in, 0, _dt(S)), S)
    with _on(S.device):
        st = lib.vvt_gram_linear_accum(
            _p(G), _p(S), _p(Z), C, N, n_out, n_in, int(with_bias), _p(ws), ws.numel(),
            _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_gram_linear_accum")
    return G


def gram_cross_linear_accum(
    X: Tensor, S: Tensor, Z: Tensor, Dl: Tensor, Zg: Tensor, with_bias: bool
) -> Tensor:
    """``X [(c,n), m] += (Z Zg^T + with_bias)[n,m] * (S Dl^T)[(c,n), m]``."""
    S, Z, Dl, Zg = _c(S), _c(Z), _c(Dl), _c(Zg)
    _chk(X, S, Z, Dl, Zg)
    C, N, n_out = S.shape
    n_in = Z.shape[1]
    n_g = Dl.shape[0]
    lib = _lib.load()
    ws = _ws(lib.vvt_gram_linear_workspace_bytes(C, N, n_out, n_in, n_g, _dt(S)), S)
    with _on(S.device):
        st = lib.vvt_gram_cross_linear_accum(
            _p(X), _p(S), _p(Z), _p(Dl), _p(Zg), C, N, n_g, n_out, n_in, int(with_bias),
            _p(ws), ws.numel(), _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_gram_cross_linear_accum")
    return X


# --------------------------------------------------------------------------
# (3) eigensolver
# --------------------------------------------------------------------------

class SyevjNotConverged(RuntimeError):
    """``vvt_syevj`` ran out of sweeps (the reference's ``Tensor.symeig`` raises ``RuntimeError`` too)."""


def syevj(G: Tensor, vectors: bool = True, return_info: bool = False):
    """Ascending eigenvalues (and eigenvectors as columns) of symmetric positive semi-definite
    ``G [R, R]`` (upper triangle read).  Raises ``SyevjNotConverged`` when the sweep limit is hit;
    ``return_info=True`` appends ``{"sweeps", "converged"}`` and leaves the decision to the caller."""
    evals, evecs, infos = _syevj_batched(G.unsqueeze(0), vectors, "vvt_syevj")
    info = infos[0]
    if return_info:
        return evals[0], (evecs[0] if vectors else None), info
    if not info["converged"]:
        raise SyevjNotConverged(f"vvt_syevj did not converge in {info['sweeps']} sweeps (R = {G.shape[0]})")
    return evals[0], (evecs[0] if vectors else None)


def syevj_batched(G: Tensor, vectors: bool = True, return_info: bool = False):
    """Eigendecompositions of ``B`` independent symmetric PSD matrices ``G [B, R, R]`` in ONE call
    (the per-group Grams of block-diagonal ``param_groups``, ``vivit/utils/hooks.py:214-219``): the batch
    index is part of every kernel's grid.  Returns ``evals [B, R]`` ascending, ``evecs [B, R, R]``
    (columns) or ``None``."""
    evals, evecs, infos = _syevj_batched(G, vectors, "vvt_syevj_batched")
    if return_info:
        return evals, evecs, infos
    if not all(i["converged"] for i in infos):
        raise SyevjNotConverged(f"vvt_syevj_batched did not converge (R = {G.shape[1]}, batch = {G.shape[0]}): {infos}")
    return evals, evecs


def syevj_dist(comm_ptr: int, world: int, G: Tensor, vectors: bool = True, return_info: bool = False,
               p2p: bool = False):
    """``syevj`` with the rounds of the two-level path distributed over the ranks of an NCCL communicator
    (``vvt_syevj_dist``): every rank passes the same ``G [R, R]`` and gets the same results.  ``comm_ptr`` is the raw
    ``ncclComm_t`` (``ProcessGroupNCCL._comm_ptr()``), ``world`` its size.  ``p2p``: the peer-memory arenas of
    exactly these ranks are mapped (``ShardedReduce.solver_arena``), so blocks change owner by direct stores over
    NVLink; otherwise by ``ncclSend`` / ``ncclRecv``.  Collective call."""
    G = _c(G)
    _chk(G)
    if G.dim() != 2 or G.shape[0] != G.shape[1]:
        raise ValueError(f"expected [R, R], got {tuple(G.shape)}")
    R = G.shape[0]
    evals = torch.empty(R, dtype=G.dtype, device=G.device)
    evecs = torch.empty(R, R, dtype=G.dtype, device=G.device) if vectors else None
    info = (ctypes.c_int * 2)(0, 1)
    if R > 0:
        lib = _lib.load()
        ws = _ws(lib.vvt_syevj_dist_workspace_bytes(R, int(vectors), _dt(G), int(world)), G)
        with _on(G.device):
            st = lib.vvt_syevj_dist(
                ctypes.c_void_p(comm_ptr), _p(evals), _p(evecs), _p(G), R, int(vectors), _p(ws), ws.numel(), info,
                int(bool(p2p)), _dt(G), _stream(G),
            )
        _lib.check(st, "vvt_syevj_dist")
    result = {"sweeps": int(info[0]), "converged": bool(info[1])}
    if return_info:
        return evals, evecs, result
    if not result["converged"]:
        raise SyevjNotConverged(f"vvt_syevj_dist did not converge in {result['sweeps']} sweeps (R = {R})")
    return evals, evecs


def _syevj_batched(G: Tensor, vectors: bool, what: str):
    G = _c(G)
    _chk(G)
    if G.dim() != 3 or G.shape[1] != G.shape[2]:
        raise ValueError(f"expected [B, R, R], got {tuple(G.shape)}")
    B, R = G.shape[0], G.shape[1]
    evals = torch.empty(B, R, dtype=G.dtype, device=G.device)
    evecs = torch.empty(B, R, R, dtype=G.dtype, device=G.device) if vectors else None
    infos = [{"sweeps": 0, "converged": True} for _ in range(B)]
    if R == 0 or B == 0:
        return evals, evecs, infos
    lib = _lib.load()
    ws = _ws(lib.vvt_syevj_batched_workspace_bytes(R, B, int(vectors), _dt(G)), G)
    info = (ctypes.c_int * (2 * B))()
    with _on(G.device):
        st = lib.vvt_syevj_batched(
            _p(evals), _p(evecs), _p(G), R, B, int(vectors), _p(ws), ws.numel(), info,
            _dt(G), _stream(G),
        )
    _lib.check(st, what)
    for b in range(B):
        infos[b] = {"sweeps": int(info[2 * b]), "converged": bool(info[2 * b + 1])}
    return evals, evecs, infos


def filter_nonzero(evals: Tensor, atol: float = 1e-7, rtol: float = 1e-5) -> Tensor:
    """Boolean mask of eigenvalues that are not ``isclose`` to zero."""
    evals = _c(evals)
    _chk(evals)
    mask = torch.empty(evals.numel(), dtype=torch.uint8, device=evals.device)
    with _on(evals.device):
        st = _lib.load().vvt_filter_nonzero(
            _p(mask), _p(evals), evals.numel(), atol, rtol, None, _dt(evals), _stream(evals)
        )
    _lib.check(st, "vvt_filter_nonzero")
    return mask.bool()


# --------------------------------------------------------------------------
# (4) back-transform, directional derivatives, Newton step
# --------------------------------------------------------------------------


def backtransform_dense(U: Tensor, V: Tensor, norm2: Optional[Tensor]) -> Tensor:
    """``E [K, D] = U [K, R] @ V [R, D]``; adds row squared norms into ``norm2`` (float64 ``[K]``)."""
    U, V = _c(U), _c(V)
    _chk(U, V)
    _chk(norm2)  # float64 accumulator whatever the compute dtype
    K, R = U.shape
    D = V.shape[1]
    E = torch.empty(K, D, dtype=V.dtype, device=V.device)
    if K == 0 or D == 0:
        return E
    with _on(V.device):
        st = _lib.load().vvt_backtransform_dense(
            _p(E), _p(norm2), _p(U), _p(V), K, R, D, _dt(V), _stream(V)
        )
    _lib.check(st, "vvt_backtransform_dense")
    return E


def backtransform_linear(U: Tensor, S: Tensor, Z: Tensor, norm2: Optional[Tensor]) -> Tensor:
    """``E[k,o,i] = sum_{c,n} U[k,c,n] S[c,n,o] Z[n,i]`` -> ``[K, out, in]``."""
    U, S, Z = _c(U), _c(S), _c(Z)
    _chk(U, S, Z)
    _chk(norm2)
    K = U.shape[0]
    C, N, n_out = S.shape
    n_in = Z.shape[1]
    E = torch.empty(K, n_out, n_in, dtype=S.dtype, device=S.device)
    if K == 0:
        return E
    ws = _ws(K * N * n_out * S.element_size(), S)
    with _on(S.device):
        st = _lib.load().vvt_backtransform_linear(
            _p(E), _p(None), _p(norm2), _p(U), _p(S), _p(Z), K, C, N, n_out, n_in,
            _p(ws), ws.numel(), _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_backtransform_linear")
    return E


def vt_mat_prod_linear(S: Tensor, Z: Tensor, M: Tensor) -> Tensor:
    """``out[f,c,n] = sum_{o,i} S[c,n,o] M[f,o,i] Z[n,i]`` -> ``[F, C, N]``."""
    S, Z, M = _c(S), _c(Z), _c(M)
    _chk(S, Z, M)
    C, N, n_out = S.shape
    n_in = Z.shape[1]
    F_ = M.shape[0]
    out = torch.empty(F_, C, N, dtype=S.dtype, device=S.device)
    ws = _ws(F_ * N * n_out * S.element_size(), S)
    with _on(S.device):
        st = _lib.load().vvt_vt_mat_prod_linear(
            _p(out), _p(S), _p(Z), _p(M), F_, C, N, n_out, n_in, _p(ws), ws.numel(),
            _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_vt_mat_prod_linear")
    return out


def scale_rows_rsqrt(E: Tensor, norm2: Tensor) -> Tensor:
    """``E[k] /= sqrt(norm2[k])`` in place (``norm2`` float64)."""
    _chk(E)
    _chk(norm2)
    K = E.shape[0]
    if E.numel() == 0:
        return E
    with _on(E.device):
        st = _lib.load().vvt_scale_rows_rsqrt(
            _p(E), _p(norm2), K, E.numel() // K, _dt(E), _stream(E)
        )
    _lib.check(st, "vvt_scale_rows_rsqrt")
    return E


def dirderiv_epilogue(
    G: Tensor, X: Tensor, U: Tensor, evals: Tensor, C: int, N_ggn: int, N: int
) -> Tuple[Tensor, Tensor]:
    """First- and second-order directional derivatives from Gram-space quantities."""
    G, X, U, evals = _c(G), _c(X), _c(U), _c(evals)
    _chk(G, X, U, evals)
    R, K = U.shape
    n_g = X.shape[1]
    gammas = torch.empty(n_g, K, dtype=G.dtype, device=G.device)
    lambdas = torch.empty(N_ggn, K, dtype=G.dtype, device=G.device)
    if K == 0:
        return gammas, lambdas
    ws = _ws(R * K * G.element_size() + 4096 + _lib.load().vvt_gram_workspace_bytes(R, K, R, _dt(G)), G)
    with _on(G.device):
        st = _lib.load().vvt_dirderiv_epilogue(
            _p(gammas), _p(lambdas), _p(G), _p(X), _p(U), _p(evals), C, N_ggn, n_g, K, N,
            _p(ws), ws.numel(), _dt(G), _stream(G),
        )
    _lib.check(st, "vvt_dirderiv_epilogue")
    return gammas, lambdas


def newton_coeff(
    U: Tensor, gammas: Tensor, lambdas: Tensor, deltas: Tensor, evals: Tensor, corr: float
) -> Tensor:
    """Gram-space Newton vector ``v [R]``."""
    U, gammas, lambdas, deltas, evals = map(_c, (U, gammas, lambdas, deltas, evals))
    _chk(U, gammas, lambdas, deltas, evals)
    R, K = U.shape
    v = torch.zeros(R, dtype=U.dtype, device=U.device)
    if K == 0:
        return v
    with _on(U.device):
        st = _lib.load().vvt_newton_coeff(
            _p(v), _p(U), _p(gammas), _p(lambdas), _p(deltas), _p(evals), R, K,
            gammas.shape[0], lambdas.shape[0], float(corr), _dt(U), _stream(U),
        )
    _lib.check(st, "vvt_newton_coeff")
    return v


def v_apply_dense(v: Tensor, V: Tensor) -> Tensor:
    """``step [D] = v [R] @ V [R, D]``."""
    v, V = _c(v), _c(V)
    _chk(v, V)
    R, D = V.shape
    step = torch.empty(D, dtype=V.dtype, device=V.device)
    with _on(V.device):
        st = _lib.load().vvt_v_apply_dense(_p(step), _p(v), _p(V), R, D, _dt(V), _stream(V))
    _lib.check(st, "vvt_v_apply_dense")
    return step


def v_apply_linear(v: Tensor, S: Tensor, Z: Tensor) -> Tensor:
    """``step[o,i] = sum_{c,n} v[c,n] S[c,n,o] Z[n,i]``."""
    v, S, Z = _c(v), _c(S), _c(Z)
    _chk(v, S, Z)
    C, N, n_out = S.shape
    n_in = Z.shape[1]
    step = torch.empty(n_out, n_in, dtype=S.dtype, device=S.device)
    ws = _ws(N * n_out * S.element_size(), S)
    with _on(S.device):
        st = _lib.load().vvt_v_apply_linear(
            _p(step), _p(None), _p(v), _p(S), _p(Z), C, N, n_out, n_in, _p(ws), ws.numel(),
            _dt(S), _stream(S),
        )
    _lib.check(st, "vvt_v_apply_linear")
    return step


# --------------------------------------------------------------------------
# per-kernel device timing (used by bench.py; off by default, no cost when off)
# --------------------------------------------------------------------------

TIMED = [
    "loss_sqrt_hessian_ce", "loss_sqrt_hessian_ce_mc", "loss_sqrt_hessian_mse", "scale_",
    "sqrt_backprop_linear", "sqrt_backprop_conv2d", "sqrt_backprop_elementwise",
    "sqrt_backprop_maxpool2d", "maxpool2d_argmax", "sqrt_backprop_avgpool2d", "v_emit_conv2d", "v_emit_bias",
    "v_emit_linear", "gemm", "gram_dense_accum", "gram_cross_accum", "gram_linear_accum",
    "gram_cross_linear_accum", "syevj", "syevj_batched", "syevj_dist", "filter_nonzero", "backtransform_dense",
    "backtransform_linear", "vt_mat_prod_linear", "scale_rows_rsqrt", "dirderiv_epilogue",
    "newton_coeff", "v_apply_dense", "v_apply_linear", "center_rows",
]


class _Timing:
    enabled = False
    records = []  # (name, start_event, end_event, [tensor-argument shapes], launches, [result shapes])


def timing_start() -> None:
    """Bracket every kernel entry point with CUDA events on the launching stream."""
    _Timing.records = []
    _Timing.enabled = True


def timing_stop():
    """Stop timing; returns ``[(name, ms, argument shapes, launches, result shapes)]`` (synchronises the device)."""
    _Timing.enabled = False
    torch.cuda.synchronize()
    out = [(n, e0.elapsed_time(e1), shp, nl, oshp) for n, e0, e1, shp, nl, oshp in _Timing.records]
    _Timing.records = []
    return out


def _timed(name, fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if not _Timing.enabled:
            return fn(*args, **kwargs)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        l0 = launch_count()
        e0.record()
        out = fn(*args, **kwargs)
        e1.record()
        shapes = [tuple(a.shape) for a in args if isinstance(a, Tensor)]
        outs = out if isinstance(out, (tuple, list)) else (out,)
        out_shapes = [tuple(o.shape) for o in outs if isinstance(o, Tensor)]
        _Timing.records.append((name, e0, e1, shapes, launch_count() - l0, out_shapes))
        return out

    return wrapper


for _name in TIMED:
    globals()[_name] = _timed(_name, globals()[_name])
del _name
