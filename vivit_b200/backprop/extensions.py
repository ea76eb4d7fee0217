"""Extensions of the native backprop engine: the sqrt-GGN factor and per-sample gradients.

Public classes mirror what the reference hands to ``with backpack(...)``:

* ``ViViTGGNExact`` / ``ViViTGGNMC`` (``vivit/extensions/secondorder/vivit/__init__.py:136-181``):
  per parameter, ``param.vivit_ggn_{exact,mc}`` = dict of closures ``gram_mat``,
  ``V_mat_prod``, ``V_t_mat_prod``;
* ``SqrtGGNExact`` / ``SqrtGGNMC`` and ``BatchGrad`` ([BackPACK], used by
  ``vivit/optim/utils.py:8-25`` and ``vivit/optim/directional_derivatives.py:127-132``):
  ``param.sqrt_ggn_{exact,mc}`` ``[C, N, *p]`` and ``param.grad_batch`` ``[N, *p]``.  With
  ``lazy=True`` (what the Computations use) the savefield holds a ``Factor`` /
  ``GradFactor`` object instead of the materialised tensor.

Module coverage: ``CrossEntropyLoss``, ``MSELoss``, ``Linear`` (also with additional input
dimensions), ``Conv2d``, ``Conv1d``, ``Conv3d`` and ``ConvTranspose1d/2d/3d`` (composed from the 2-d kernels,
``backprop/conv_nd.py``), ``BatchNorm1d/2d/3d`` (evaluation mode), ``ReLU``, ``Sigmoid``, ``Tanh``,
``LeakyReLU``, ``ELU``, ``SELU``, ``LogSigmoid``, ``MaxPool2d``, ``AvgPool2d``, ``MaxPool1d``, ``AvgPool1d``
(1-d layers run on the 2-d kernels over feature maps of unit height), ``ZeroPad2d``,
``Flatten``, ``Dropout``, ``Identity``, and the branching modules of ``vivit_b200.custom_module``
(``Parallel`` / ``SumModule``, ``ScaleModule``, ``Pad``, ``Slicing``).
Anything else raises ``NotImplementedError`` (the reference's ``fail_mode="ERROR"``,
``__init__.py:83``).
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from vivit_b200 import kernels
from vivit_b200.backprop import conv_nd
from vivit_b200.custom_module import Pad, ScaleModule, Slicing, SumModule
from vivit_b200.factors import (
    DenseFactor,
    DenseGrad,
    Factor,
    LinearBiasFactor,
    LinearWeightFactor,
    LinearWeightGrad,
    StreamedConvFactor,
    conv_factor_streaming,
    link_linear_factors,
)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Extension:
    """Base class: dispatches a module to its handler and owns a savefield name."""

    savefield: str = ""

    def __init__(self, subsampling: Optional[List[int]] = None):
        self._subsampling = None if subsampling is None else list(subsampling)
        self._sub_cache: Dict[torch.device, Tensor] = {}
        self._shard = None  # (rank, world): this rank owns a dim-0 slice of every parameter

    def _own(self, size: int):
        """``[lo, hi)`` of a parameter's leading dimension owned by this rank (``vivit_b200.dist``)."""
        if self._shard is None:
            return 0, size
        from vivit_b200.dist import shard_bounds

        return shard_bounds(size, *self._shard)

    def get_subsampling(self) -> Optional[List[int]]:
        return self._subsampling

    def _sub_index(self, device) -> Optional[Tensor]:
        if self._subsampling is None:
            return None
        if device not in self._sub_cache:
            self._sub_cache[device] = torch.tensor(self._subsampling, dtype=torch.int64, device=device)
        return self._sub_cache[device]

    def _subsample(self, t: Tensor) -> Tensor:
        """``backpack.utils.subsampling.subsample`` along the batch axis."""
        idx = self._sub_index(t.device)
        return t if idx is None else t.index_select(0, idx)

    def _begin(self) -> None:  # called when a backpack context is entered
        pass

    def _apply(self, module: nn.Module, g_out: Tensor) -> None:
        raise NotImplementedError

    @staticmethod
    def _unsupported(ext, module):
        raise NotImplementedError(f"Extension {type(ext).__name__} does not support {type(module)}")


def _trainable(module: nn.Module, name: str):
    p = getattr(module, name, None)
    return p if isinstance(p, nn.Parameter) and p.requires_grad else None


# --------------------------------------------------------------------------
# second-order: symmetric factor of the GGN
# --------------------------------------------------------------------------


class _SqrtFactorExtension(Extension):
    """Back-propagates ``S`` (``[V, N_sub, *features]``) and leaves a factor per parameter."""

    def __init__(self, savefield: str, subsampling, mc_samples: int, lazy: bool, closures: bool):
        super().__init__(subsampling)
        self.savefield = savefield
        self._field = "_vvt_S_" + savefield  # attribute carrying S on activations
        self._mc_samples = mc_samples
        self._lazy = lazy
        self._closures = closures
        self.mc_state: Optional[Tensor] = None  # tests may pin the random draw here

    def get_loss_hessian_strategy(self) -> str:
        return "sampling" if self._mc_samples else "exact"

    def get_num_mc_samples(self) -> int:
        return self._mc_samples

    # ---- engine entry ---------------------------------------------------------
    def _apply(self, module: nn.Module, g_out: Tensor) -> None:
        if isinstance(module, (nn.CrossEntropyLoss, nn.MSELoss)):
            S = self._loss_factor(module)
            setattr(module.input0, self._field, S)
            return
        out = module.output
        S = getattr(out, self._field, None)
        if S is None:
            raise RuntimeError(
                f"{type(self).__name__}: no back-propagated factor at the output of {module}. "
                "Extend the loss function too and call it on the model output."
            )
        delattr(out, self._field)
        if hasattr(out, self._field + "_sum"):
            delattr(out, self._field + "_sum")
        handler = _FACTOR_HANDLERS.get(type(module))
        if handler is None:
            for cls, h in _FACTOR_HANDLERS.items():
                if isinstance(module, cls):
                    handler = h
                    break
        if handler is None:
            self._unsupported(self, module)
        if handler is _factor_sum:  # the only handler with several inputs
            for idx in range(_num_inputs(module)):
                inp = getattr(module, f"input{idx}")
                if isinstance(inp, Tensor) and inp.requires_grad:
                    self._pass_on(inp, S)
            return
        need_in = isinstance(module.input0, Tensor) and module.input0.requires_grad
        S_in = handler(self, module, S, need_in)
        if need_in and S_in is not None:
            self._pass_on(module.input0, S_in)

    def _pass_on(self, inp: Tensor, S_in: Tensor) -> None:
        """Attach the factor to the tensor it belongs to.  A tensor that feeds several modules collects
        the sum (``accumulate_backpropagated_quantities``, ``secondorder/vivit/__init__.py:130-133``)."""
        existing = getattr(inp, self._field, None)
        if existing is None:
            setattr(inp, self._field, S_in)
            return
        if existing.shape != S_in.shape:
            raise RuntimeError(f"branches disagree on the factor shape: {existing.shape} vs {S_in.shape}")
        # the first sum goes into a fresh buffer: SumModule / Identity / Flatten hand the SAME storage to
        # several tensors, which must not see each other's contributions
        if not getattr(inp, self._field + "_sum", False):
            existing = existing.clone(memory_format=torch.contiguous_format)
            setattr(inp, self._field, existing)
            setattr(inp, self._field + "_sum", True)
        kernels.axpy_(existing, S_in.contiguous())

    # ---- loss ---------------------------------------------------------------------
    def _loss_factor(self, module) -> Tensor:
        out = module.input0.detach()
        mean = module.reduction == "mean"
        if module.reduction not in ("mean", "sum"):
            raise NotImplementedError("loss reduction must be 'mean' or 'sum'")
        if out.dim() != 2:
            raise NotImplementedError("loss factors support [N, C] model outputs only")
        sub = self._sub_index(out.device)
        n_total, C = out.shape
        n_sub = n_total if sub is None else sub.numel()
        M = self._mc_samples
        if isinstance(module, nn.CrossEntropyLoss):
            if M == 0:
                return kernels.loss_sqrt_hessian_ce(out, sub, mean)
            ids = self.mc_state
            if ids is None:
                # the one random draw of the MC factor, with torch's generator as in [BackPACK]
                probs = self._subsample(F.softmax(out, dim=1))
                ids = torch.multinomial(probs, M, replacement=True).t().contiguous()
            return kernels.loss_sqrt_hessian_ce_mc(out, sub, ids.to(out.device), mean)
        scale = math.sqrt(2.0) / (math.sqrt(out.numel()) if mean else 1.0)
        if M == 0:
            return kernels.loss_sqrt_hessian_mse(n_sub, C, scale, out)
        normal = self.mc_state
        if normal is None:
            normal = torch.randn(M, n_sub, C, dtype=out.dtype, device=out.device)
        return kernels.scale_(normal.to(out).clone(), scale / math.sqrt(M))

    # ---- parameters ---------------------------------------------------------------
    def _save(self, param: nn.Parameter, factor: Factor) -> None:
        if self._closures:
            value = {
                "V_mat_prod": lambda mat, f=factor: f.backtransform(mat.reshape(mat.shape[0], -1), None),
                "V_t_mat_prod": lambda mat, f=factor: f.vt_mat_prod(mat),
                "gram_mat": lambda f=factor: f.gram_mat(),
                "_factor": factor,
            }
        elif self._lazy:
            value = factor
        else:
            value = factor.materialize()
        setattr(param, self.savefield, value)


def _linear_as_conv(S: Tensor, z: Tensor):
    """A Linear layer whose input has additional dimensions ``[N, *E, in]`` acts as a 1x1 convolution over
    the ``E`` extra positions: re-lay ``S [V, N, *E, out]`` and ``z [N, *E, in]`` as the channel-major maps
    ``[V, N, out, E, 1]`` / ``[N, in, E, 1]`` the conv emit kernels read (index plumbing only).  The reference
    materialises the factor of such a layer through ``param_mjp`` (``linear.py:26-27,38-39``)."""
    V, N = S.shape[:2]
    n_out, n_in = S.shape[-1], z.shape[-1]
    E = z.numel() // max(N * n_in, 1)
    Sc = S.reshape(V, N, E, n_out).transpose(2, 3).contiguous().reshape(V, N, n_out, E, 1)
    Xc = z.reshape(N, E, n_in).transpose(1, 2).contiguous().reshape(N, n_in, E, 1)
    return Sc, Xc


_ONE = (1, 1)


def _emit_linear_extra(Sc: Tensor, Xc: Tensor) -> Tensor:
    """``[V, N, out, in]``: sum over the extra positions of ``S (x) z`` (``einsum("vn...o,n...i->vnoi")``)."""
    Vt = kernels.v_emit_conv2d(Sc, Xc, _ONE, _ONE, (0, 0), _ONE)
    return Vt.reshape(*Vt.shape[:4])


def _factor_linear(ext: _SqrtFactorExtension, module: nn.Linear, S: Tensor, need_in: bool):
    z = ext._subsample(module.input0.detach())
    w, b = _trainable(module, "weight"), _trainable(module, "bias")
    lo, hi = ext._own(S.shape[-1])
    S_own = S if hi - lo == S.shape[-1] else S[..., lo:hi].contiguous()
    if z.dim() > 2:  # additional dimensions: dense factors
        if w is not None or b is not None:
            Sc, Xc = _linear_as_conv(S_own, z)
            if b is not None:
                ext._save(b, DenseFactor(kernels.v_emit_bias(Sc), (hi - lo,)))
            if w is not None:
                ext._save(w, DenseFactor(_emit_linear_extra(Sc, Xc), (hi - lo, z.shape[-1])))
    else:
        bf = LinearBiasFactor(S_own, (hi - lo,)) if b is not None else None
        wf = LinearWeightFactor(S_own, z.contiguous()) if w is not None else None
        if bf is not None and wf is not None:
            link_linear_factors(wf, w, bf, b)  # one group => one structured Gram call for both (factors.py)
        if bf is not None:  # bias first, as [BackPACK] does (params=["bias", "weight"], linear.py:24)
            ext._save(b, bf)
        if wf is not None:
            ext._save(w, wf)
    return kernels.sqrt_backprop_linear(S, module.weight.detach()) if need_in else None


def _conv_check(module):
    if module.groups != 1:
        raise NotImplementedError("grouped convolutions are not supported")
    if module.padding_mode != "zeros" or isinstance(module.padding, str):
        raise NotImplementedError("convolutions need numeric zero padding")


def _conv_geometry(module):
    """``(kernel, stride, padding, dilation)`` as 2-d pairs; a 1-d convolution is a 2-d one of unit height."""
    if isinstance(module, nn.Conv1d):
        return ((1, module.kernel_size[0]), (1, module.stride[0]), (0, module.padding[0]), (1, module.dilation[0]))
    return (_pair(module.kernel_size), _pair(module.stride), _pair(module.padding), _pair(module.dilation))


def _factor_conv2d(ext: _SqrtFactorExtension, module: nn.Conv2d, S: Tensor, need_in: bool):
    _conv_check(module)
    one_d = isinstance(module, nn.Conv1d)
    x = ext._subsample(module.input0.detach())
    if one_d:  # [V, N, Co, L'] / [N, Ci, L] -> unit-height feature maps
        S, x = S.unsqueeze(3), x.unsqueeze(2)
    w, b = _trainable(module, "weight"), _trainable(module, "bias")
    kernel, *geom = _conv_geometry(module)
    lo, hi = ext._own(S.shape[2])
    S_own = S if hi - lo == S.shape[2] else S[:, :, lo:hi].contiguous()
    if b is not None:
        ext._save(b, DenseFactor(kernels.v_emit_bias(S_own), (hi - lo,)))
    if w is not None:
        w_shape = (hi - lo, *w.shape[1:])
        stream = conv_factor_streaming()
        if stream is not None:  # opt-in: V_p^T emitted chunk by chunk at every use, never materialised as a whole
            ext._save(w, StreamedConvFactor(S_own, x, kernel, geom, w_shape, stream))
        else:
            Vt = kernels.v_emit_conv2d(S_own, x, kernel, *geom)
            ext._save(w, DenseFactor(Vt, w_shape))
    if not need_in:
        return None
    weight = module.weight.detach()
    S_in = kernels.sqrt_backprop_conv2d(S, weight.unsqueeze(2) if one_d else weight, tuple(x.shape[2:]), *geom)
    return S_in.squeeze(3) if one_d else S_in


_CONV_TRANSPOSE = (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)


def _conv3d_operands(module, S: Tensor, x: Tensor):
    """Factor, input, weight and geometry of a ``Conv3d`` / ``ConvTransposeNd`` layer in the 3-d form of
    ``backprop/conv_nd.py`` (1-d and 2-d layers get unit depth / height)."""
    nd = x.dim() - 2
    geom = (
        conv_nd.triple(module.kernel_size, nd, 1),
        conv_nd.triple(module.stride, nd, 1),
        conv_nd.triple(module.padding, nd, 0),
        conv_nd.triple(module.dilation, nd, 1),
    )
    return conv_nd.lift_to_3d(S, nd), conv_nd.lift_to_3d(x, nd), conv_nd.lift_to_3d(module.weight.detach(), nd), nd, geom


def _unlift(t: Tensor, nd: int) -> Tensor:
    """Inverse of ``conv_nd.lift_to_3d`` on ``[V, N, C, D, H, W]``."""
    for _ in range(3 - nd):
        t = t.squeeze(3)
    return t


def _factor_conv3d(ext: _SqrtFactorExtension, module: nn.Conv3d, S: Tensor, need_in: bool):
    """``Conv3d`` as a sum of 2-d problems over the depth taps (``backprop/conv_nd.py``)."""
    _conv_check(module)
    S, x, weight, nd, (kernel, *geom) = _conv3d_operands(module, S, ext._subsample(module.input0.detach()))
    w, b = _trainable(module, "weight"), _trainable(module, "bias")
    lo, hi = ext._own(S.shape[2])
    S_own = S if hi - lo == S.shape[2] else S[:, :, lo:hi].contiguous()
    if b is not None:
        ext._save(b, DenseFactor(kernels.v_emit_bias(S_own), (hi - lo,)))
    if w is not None:
        ext._save(w, DenseFactor(conv_nd.conv3d_emit(S_own, x, kernel, *geom), (hi - lo, *w.shape[1:])))
    return conv_nd.conv3d_backprop(S, weight, tuple(x.shape[2:]), *geom) if need_in else None


def _factor_conv_transpose(ext: _SqrtFactorExtension, module, S: Tensor, need_in: bool):
    """``ConvTranspose1d/2d/3d``: the 2-d kernels with the roles of their operands swapped
    (``backprop/conv_nd.py``).  The weight is ``[C_in, C_out, *k]``: a rank owns a slice of ``C_in`` of it and
    a slice of ``C_out`` of the bias."""
    _conv_check(module)
    S, x, weight, nd, (kernel, *geom) = _conv3d_operands(module, S, ext._subsample(module.input0.detach()))
    w, b = _trainable(module, "weight"), _trainable(module, "bias")
    if b is not None:
        lo, hi = ext._own(S.shape[2])
        ext._save(b, DenseFactor(kernels.v_emit_bias(S[:, :, lo:hi].contiguous()), (hi - lo,)))
    if w is not None:
        lo, hi = ext._own(x.shape[1])
        Vt = conv_nd.conv_transpose3d_emit(S, x[:, lo:hi].contiguous(), kernel, *geom)
        ext._save(w, DenseFactor(Vt, (hi - lo, *w.shape[1:])))
    if not need_in:
        return None
    return _unlift(conv_nd.conv_transpose3d_backprop(S, weight, tuple(x.shape[2:]), *geom), nd)


def _factor_act(act, use_output, scale_of=None):
    def handler(ext, module, S, need_in):
        if not need_in:
            return None
        ref = module.output if use_output else module.input0
        scale = 1.0 if scale_of is None else float(scale_of(module))
        return kernels.sqrt_backprop_elementwise(S, ext._subsample(ref.detach()), act, scale)

    return handler


_BATCHNORM = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)


def _bn_check(module) -> None:
    # [BackPACK] mixes the samples of a batch in training mode; the reference's fixtures and the GGN
    # factorisation per sample only make sense in evaluation mode (test/settings.py:118-160)
    if module.training or not module.track_running_stats or module.weight is None:
        raise NotImplementedError("BatchNorm is supported in evaluation mode (affine, running statistics) only")


def _bn_normalized_input(module, x: Tensor) -> Tensor:
    """``(x - running_mean) / sqrt(running_var + eps)``: a quantity of the user's forward pass, recomputed
    with torch's own batch-norm op (as the max-pool handler recomputes the arg-max positions)."""
    return F.batch_norm(x, module.running_mean, module.running_var, None, None, False, 0.0, module.eps)


def _bn_param_rows(S: Tensor, x_hat: Tensor):
    """Per-sample factors of the BatchNorm parameters from ``S [V, N, C, *spatial]``: bias ``sum_x S``,
    weight ``sum_x S * x_hat`` ([BackPACK] ``BatchNormNdDerivatives`` in evaluation mode)."""
    V, N, C = S.shape[:3]
    S3 = S.reshape(V, N, C, -1)
    prod = kernels.sqrt_backprop_elementwise(S3, x_hat, kernels.ACT_MUL)
    if S3.shape[3] == 1:
        return S3.reshape(V, N, C), prod.reshape(V, N, C)
    return kernels.v_emit_bias(S3), kernels.v_emit_bias(prod)


def _factor_batchnorm(ext, module, S, need_in):
    _bn_check(module)
    w, b = _trainable(module, "weight"), _trainable(module, "bias")
    C = S.shape[2]
    lo, hi = ext._own(C)
    if w is not None or b is not None:
        x_hat = _bn_normalized_input(module, ext._subsample(module.input0.detach()))
        S_own = S if hi - lo == C else S[:, :, lo:hi].contiguous()
        Vb, Vw = _bn_param_rows(S_own, x_hat if hi - lo == C else x_hat[:, lo:hi].contiguous())
        if b is not None:
            ext._save(b, DenseFactor(Vb, (hi - lo,)))
        if w is not None:
            ext._save(w, DenseFactor(Vw, (hi - lo,)))
    if not need_in:
        return None
    # Jacobian w.r.t. the input: one factor per channel, laid out over one sample's features
    per_channel = module.weight.detach() / (module.running_var + module.eps).sqrt()
    spatial = S[0, 0, 0].numel()
    ref = per_channel.to(S.dtype).reshape(C, 1).expand(C, spatial).contiguous()
    return kernels.sqrt_backprop_elementwise(S, ref, kernels.ACT_MUL)


def _factor_zeropad2d(ext, module: nn.ZeroPad2d, S, need_in):
    """Crop: the Jacobian of zero padding drops the padded border ([BackPACK] ``ZeroPad2dDerivatives``)."""
    left, right, top, bottom = module.padding
    h, w = S.shape[-2:]
    return S[..., top : h - bottom, left : w - right].contiguous()


def _factor_pad(ext, module: Pad, S, need_in):
    """Crop: constant padding has no derivative on the padded border, whatever the fill value
    ([BackPACK] ``PadDerivatives``; fixture ``test/settings.py:167``)."""
    if module.mode != "constant":
        raise NotImplementedError("Pad is supported with mode='constant' only")
    pad = module.pad
    if len(pad) % 2 or len(pad) // 2 > S.dim() - 2:
        raise NotImplementedError("Pad must pad feature dimensions only (not the batch axis)")
    index = [slice(None)] * S.dim()
    for i in range(len(pad) // 2):
        axis = S.dim() - 1 - i
        index[axis] = slice(pad[2 * i], S.shape[axis] - pad[2 * i + 1])
    return S[tuple(index)].contiguous()


def _factor_slicing(ext, module: Slicing, S, need_in):
    """Embed in zeros: entries of the input that the slice drops do not reach the output
    ([BackPACK] ``SlicingDerivatives``; fixture ``test/settings.py:171``)."""
    info = module.slice_info
    if len(info) == 0:
        return S
    if info[0] != slice(None):
        raise NotImplementedError("Slicing must keep the batch axis whole")
    x = module.input0
    S_in = torch.zeros(S.shape[0], S.shape[1], *x.shape[1:], dtype=S.dtype, device=S.device)
    S_in[(slice(None), slice(None)) + tuple(info[1:])] = S
    return S_in


def _factor_scale(ext, module: ScaleModule, S, need_in):
    """``d (w x) / dx = w I`` ([BackPACK] ``ScaleModuleDerivatives``)."""
    if not need_in or module.weight == 1.0:
        return S if need_in else None
    return kernels.scale_(S.clone(memory_format=torch.contiguous_format), module.weight)


def _factor_sum(ext, module, S, need_in):  # dispatched in _apply: every summand receives S
    return S


def _num_inputs(module) -> int:
    n = 0
    while hasattr(module, f"input{n}"):
        n += 1
    return n


def _factor_dropout(ext, module: nn.Dropout, S, need_in):
    if not module.training or module.p == 0.0:
        return S
    ref = ext._subsample(module.output.detach())
    return kernels.sqrt_backprop_elementwise(S, ref, kernels.ACT_DROPOUT, 1.0 / (1.0 - module.p))


def _factor_flatten(ext, module, S, need_in):
    x = module.input0
    return S.reshape(S.shape[0], S.shape[1], *x.shape[1:])


def _factor_identity(ext, module, S, need_in):
    return S


def _factor_maxpool2d(ext, module: nn.MaxPool2d, S, need_in):
    if not need_in:
        return None
    x = ext._subsample(module.input0.detach())
    k, s = _pair(module.kernel_size), _pair(module.stride if module.stride is not None else module.kernel_size)
    p, d = _pair(module.padding), _pair(module.dilation)
    # arg-max positions of the forward pass (index plumbing; [BackPACK] recomputes them from the layer input too)
    idx = kernels.maxpool2d_argmax(x, k, s, p, d, module.ceil_mode)
    return kernels.sqrt_backprop_maxpool2d(S, idx, tuple(x.shape[2:]), k, s, p, d)


def _factor_avgpool2d(ext, module: nn.AvgPool2d, S, need_in):
    if not need_in:
        return None
    if module.ceil_mode or not module.count_include_pad or module.divisor_override is not None:
        raise NotImplementedError("AvgPool2d: ceil_mode / count_include_pad=False / divisor_override")
    x = module.input0
    k = _pair(module.kernel_size)
    s = _pair(module.stride if module.stride is not None else module.kernel_size)
    return kernels.sqrt_backprop_avgpool2d(S, tuple(x.shape[2:]), k, s, _pair(module.padding))


def _one(v) -> int:
    return v if isinstance(v, int) else v[0]


def _factor_maxpool1d(ext, module: nn.MaxPool1d, S, need_in):
    """A 1-d pooling layer is the 2-d one over feature maps of unit height."""
    if not need_in:
        return None
    x = ext._subsample(module.input0.detach()).unsqueeze(2)
    stride = module.stride if module.stride is not None else module.kernel_size
    k, s = (1, _one(module.kernel_size)), (1, _one(stride))
    p, d = (0, _one(module.padding)), (1, _one(module.dilation))
    idx = kernels.maxpool2d_argmax(x, k, s, p, d, module.ceil_mode)
    return kernels.sqrt_backprop_maxpool2d(S.unsqueeze(3), idx, tuple(x.shape[2:]), k, s, p, d).squeeze(3)


def _factor_avgpool1d(ext, module: nn.AvgPool1d, S, need_in):
    if not need_in:
        return None
    if module.ceil_mode or not module.count_include_pad:
        raise NotImplementedError("AvgPool1d: ceil_mode / count_include_pad=False")
    stride = module.stride if module.stride is not None else module.kernel_size
    k, s, p = (1, _one(module.kernel_size)), (1, _one(stride)), (0, _one(module.padding))
    length = module.input0.shape[2]
    return kernels.sqrt_backprop_avgpool2d(S.unsqueeze(3), (1, length), k, s, p).squeeze(3)


_FACTOR_HANDLERS = {
    nn.Linear: _factor_linear,
    nn.Conv2d: _factor_conv2d,
    nn.Conv1d: _factor_conv2d,
    nn.Conv3d: _factor_conv3d,
    nn.ConvTranspose1d: _factor_conv_transpose,
    nn.ConvTranspose2d: _factor_conv_transpose,
    nn.ConvTranspose3d: _factor_conv_transpose,
    nn.ReLU: _factor_act(kernels.ACT_RELU, use_output=False),
    nn.Sigmoid: _factor_act(kernels.ACT_SIGMOID, use_output=True),
    nn.Tanh: _factor_act(kernels.ACT_TANH, use_output=True),
    nn.LeakyReLU: _factor_act(kernels.ACT_LEAKY_RELU, use_output=False, scale_of=lambda m: m.negative_slope),
    nn.ELU: _factor_act(kernels.ACT_ELU, use_output=False, scale_of=lambda m: m.alpha),
    nn.SELU: _factor_act(kernels.ACT_SELU, use_output=False),
    nn.LogSigmoid: _factor_act(kernels.ACT_LOGSIGMOID, use_output=False),
    nn.ZeroPad2d: _factor_zeropad2d,
    nn.BatchNorm1d: _factor_batchnorm,
    nn.BatchNorm2d: _factor_batchnorm,
    nn.BatchNorm3d: _factor_batchnorm,
    nn.Dropout: _factor_dropout,
    nn.Flatten: _factor_flatten,
    nn.Identity: _factor_identity,
    nn.MaxPool2d: _factor_maxpool2d,
    nn.AvgPool2d: _factor_avgpool2d,
    nn.MaxPool1d: _factor_maxpool1d,
    nn.AvgPool1d: _factor_avgpool1d,
    Pad: _factor_pad,
    ScaleModule: _factor_scale,
    Slicing: _factor_slicing,
    SumModule: _factor_sum,
}


class ViViTGGNExact(_SqrtFactorExtension):
    """Functional access to the exact GGN factor (``__init__.py:136-152``)."""

    def __init__(self, subsampling: List[int] = None):
        super().__init__("vivit_ggn_exact", subsampling, 0, lazy=True, closures=True)


class ViViTGGNMC(_SqrtFactorExtension):
    """Functional access to the MC-sampled GGN factor (``__init__.py:155-181``)."""

    def __init__(self, mc_samples: int = 1, subsampling: List[int] = None):
        super().__init__("vivit_ggn_mc", subsampling, mc_samples, lazy=True, closures=True)


class SqrtGGNExact(_SqrtFactorExtension):
    """[BackPACK] ``SqrtGGNExact``: ``param.sqrt_ggn_exact`` of shape ``[C, N, *param.shape]``."""

    def __init__(self, subsampling: List[int] = None, lazy: bool = False):
        super().__init__("sqrt_ggn_exact", subsampling, 0, lazy=lazy, closures=False)


class SqrtGGNMC(_SqrtFactorExtension):
    """[BackPACK] ``SqrtGGNMC``: ``param.sqrt_ggn_mc`` of shape ``[M, N, *param.shape]``."""

    def __init__(self, mc_samples: int = 1, subsampling: List[int] = None, lazy: bool = False):
        super().__init__("sqrt_ggn_mc", subsampling, mc_samples, lazy=lazy, closures=False)


_FACTOR_FAMILIES = {"vivit": (ViViTGGNExact, ViViTGGNMC), "sqrt_ggn": (SqrtGGNExact, SqrtGGNMC)}


def factor_extension(family: str, subsampling: Optional[List[int]], mc_samples: int, **kwargs):
    """The exact member of an extension family when ``mc_samples`` is zero, its Monte-Carlo member
    otherwise (the switch of ``vivit/linalg/utils.py:11-28`` and ``vivit/optim/utils.py:8-25``)."""
    exact, sampled = _FACTOR_FAMILIES[family]
    if mc_samples:
        return sampled(mc_samples=mc_samples, subsampling=subsampling, **kwargs)
    return exact(subsampling=subsampling, **kwargs)


# --------------------------------------------------------------------------
# first-order: per-sample gradients
# --------------------------------------------------------------------------


class BatchGrad(Extension):
    """[BackPACK] ``BatchGrad``: ``param.grad_batch[n] = d loss / d param`` of sample ``n``
    (carries the ``1/N`` of a mean reduction), optionally sub-sampled."""

    savefield = "grad_batch"

    def __init__(self, subsampling: List[int] = None, lazy: bool = False):
        super().__init__(subsampling)
        self._lazy = lazy

    def _save(self, param, grad):
        setattr(param, self.savefield, grad if self._lazy else grad.materialize())

    def _apply(self, module: nn.Module, g_out: Tensor) -> None:
        if isinstance(module, nn.Linear):
            w, b = _trainable(module, "weight"), _trainable(module, "bias")
            if w is None and b is None:
                return
            g = self._subsample(g_out.detach())
            lo, hi = self._own(g.shape[-1])
            if g.dim() > 2:  # Linear with additional input dimensions
                Sc, Xc = _linear_as_conv(g[..., lo:hi][None], self._subsample(module.input0.detach()))
                if b is not None:
                    self._save(b, DenseGrad(kernels.v_emit_bias(Sc)[0], (hi - lo,)))
                if w is not None:
                    self._save(w, DenseGrad(_emit_linear_extra(Sc, Xc)[0], (hi - lo, w.shape[1])))
                return
            g = g[:, lo:hi].contiguous()
            if not self._lazy and g.data_ptr() == g_out.data_ptr():
                # the materialised bias gradient is handed to the user (hooks may centre it in
                # place): it must not alias the gradient autograd is still back-propagating
                g = g.clone()
            z = self._subsample(module.input0.detach())
            if b is not None:
                self._save(b, DenseGrad(g, (hi - lo,)))
            if w is not None:
                self._save(w, LinearWeightGrad(g, z.contiguous()))
        elif isinstance(module, (nn.Conv2d, nn.Conv1d)):
            w, b = _trainable(module, "weight"), _trainable(module, "bias")
            if w is None and b is None:
                return
            _conv_check(module)
            g = self._subsample(g_out.detach())
            lo, hi = self._own(g.shape[1])
            g = g[:, lo:hi].contiguous()[None]  # one "class": [1, N, Co, Ho, Wo]
            x = self._subsample(module.input0.detach())
            if isinstance(module, nn.Conv1d):  # unit-height feature maps
                g, x = g.unsqueeze(3), x.unsqueeze(2)
            if b is not None:
                self._save(b, DenseGrad(kernels.v_emit_bias(g)[0], (hi - lo,)))
            if w is not None:
                gw = kernels.v_emit_conv2d(g, x, *_conv_geometry(module))[0]
                self._save(w, DenseGrad(gw, (hi - lo, *w.shape[1:])))
        elif isinstance(module, (nn.Conv3d, *_CONV_TRANSPOSE)):
            w, b = _trainable(module, "weight"), _trainable(module, "bias")
            if w is None and b is None:
                return
            _conv_check(module)
            g, x, _, _, (kernel, *geom) = _conv3d_operands(
                module, self._subsample(g_out.detach())[None], self._subsample(module.input0.detach())
            )  # one "class": [1, N, Co, D', H', W']
            transposed = isinstance(module, _CONV_TRANSPOSE)
            if b is not None:
                lo, hi = self._own(g.shape[2])
                self._save(b, DenseGrad(kernels.v_emit_bias(g[:, :, lo:hi].contiguous())[0], (hi - lo,)))
            if w is not None and transposed:
                lo, hi = self._own(x.shape[1])
                gw = conv_nd.conv_transpose3d_emit(g, x[:, lo:hi].contiguous(), kernel, *geom)[0]
                self._save(w, DenseGrad(gw, (hi - lo, *w.shape[1:])))
            elif w is not None:
                lo, hi = self._own(g.shape[2])
                gw = conv_nd.conv3d_emit(g[:, :, lo:hi].contiguous(), x, kernel, *geom)[0]
                self._save(w, DenseGrad(gw, (hi - lo, *w.shape[1:])))
        elif isinstance(module, _BATCHNORM):
            w, b = _trainable(module, "weight"), _trainable(module, "bias")
            if w is None and b is None:
                return
            _bn_check(module)
            g = self._subsample(g_out.detach())
            lo, hi = self._own(g.shape[1])
            x_hat = _bn_normalized_input(module, self._subsample(module.input0.detach()))
            gb, gw = _bn_param_rows(g[:, lo:hi].contiguous()[None], x_hat[:, lo:hi].contiguous())
            if b is not None:
                self._save(b, DenseGrad(gb[0].clone() if gb.data_ptr() == g_out.data_ptr() else gb[0], (hi - lo,)))
            if w is not None:
                self._save(w, DenseGrad(gw[0], (hi - lo,)))
        elif any(p.requires_grad for p in module.parameters(recurse=False)):
            self._unsupported(self, module)
