"""Plain-torch CPU restatement of the reference's low-rank GGN hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity is pinned against
outputs of the reference's own Computations executed in the build container
(``tests/golden/reference_run.pt``, made by ``tests/golden/make_reference_run.py``) and by
relation to the autograd GGN (``oracle/autograd_ggn.py``), the way the reference's own
tests pin it; there are no golden files upstream.

Every function names the reference lines (relative to ``/root/reference``) it
follows.  BackPACK (``backpack-for-pytorch>=1.5,<2``, ``setup.cfg:36``) is an
un-vendored dependency; the parts of it this path relies on are restated from
their published behaviour and marked ``[external]``:

* ``CrossEntropyLossDerivatives.sqrt_hessian[_sampled]``,
  ``MSELossDerivatives.sqrt_hessian[_sampled]`` -- symmetric factor of the
  loss Hessian, shape ``[C or M, N, C]``.
* ``*Derivatives.jac_t_mat_prod`` -- transposed-Jacobian product of a layer,
  applied to a ``[V, N, *out]`` stack.
* ``*Derivatives.param_mjp(..., sum_batch=False)`` -- per-sample
  transposed-Jacobian product w.r.t. a parameter, ``[V, N, *param.shape]``.
* ``BatchGrad`` -- ``grad_batch[n] = d(loss)/d(param)`` restricted to sample
  ``n`` (includes the ``1/N`` of a mean reduction).

``Tensor.symeig(eigenvectors, upper=True)`` is replaced by
``torch.linalg.eigh(UPLO="U")`` (same ordering and column convention).

The model must be a (possibly nested) ``torch.nn.Sequential`` of supported leaf
modules and branching containers (``Parallel``: every branch sees the same input,
the outputs are summed); the reference's hook machinery (``vivit/utils/hooks.py``)
is replaced by an explicit reverse loop, which visits parameters in the same order.
"""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, einsum, nn

# --------------------------------------------------------------------------
# model plumbing
# --------------------------------------------------------------------------


def _is_parallel(model: nn.Module) -> bool:
    """A branching container ([BackPACK] ``custom_module.branching.Parallel`` [external]; restated for the
    tests in ``vivit_b200.custom_module``): duck-typed so that the oracle imports nothing of the product."""
    return hasattr(model, "branches") and hasattr(model, "merge_module")


def forward_capture(model: nn.Module, x: Tensor, retain_grad: bool = False):
    """Run the model, keeping every leaf's input and output (BackPACK's ``module.input0`` /
    ``module.output`` [external]).  Records are nested like the model: ``("leaf", module, input,
    output)`` in execution order, and ``("parallel", [records of each branch], input, output)`` for a
    branching container whose branch outputs are summed."""
    if isinstance(model, nn.Sequential):
        records = []
        for child in model.children():
            recs, x = forward_capture(child, x, retain_grad)
            records += recs
        return records, x
    if _is_parallel(model):
        branches, outs = [], []
        for branch in model.branches():
            recs, o = forward_capture(branch, x, retain_grad)
            branches.append(recs)
            outs.append(o)
        y = outs[0]
        for o in outs[1:]:
            y = y + o
        return [("parallel", branches, x, y)], y
    y = model(x)
    if retain_grad and y.requires_grad:  # false up to the first layer with parameters
        y.retain_grad()
    return [("leaf", model, x, y)], y


def _sub(t: Tensor, subsampling: Optional[Sequence[int]]) -> Tensor:
    """``backpack.utils.subsampling.subsample`` [external]."""
    return t if subsampling is None else t[list(subsampling)]


# --------------------------------------------------------------------------
# a1-a3: symmetric factors of the loss Hessian  [external]
# --------------------------------------------------------------------------


def sqrt_hessian_ce(
    logits: Tensor, subsampling: Optional[Sequence[int]], reduction: str = "mean"
) -> Tensor:
    """``S[v,n,c] = tau_nc (delta_vc - tau_nv tau_nc) / sqrt(N)``, ``tau = sqrt(softmax)``.

    ``N`` is the *full* batch size even under sub-sampling (wired at
    ``vivit/extensions/secondorder/vivit/__init__.py:86``).
    """
    n_total = logits.shape[0]
    probs = _sub(F.softmax(logits, dim=1), subsampling)
    tau = probs.sqrt()
    c = probs.shape[1]
    eye = torch.eye(c, dtype=logits.dtype, device=logits.device)
    s = einsum("nc,vnc->vnc", tau, eye[:, None, :] - einsum("nv,nc->vnc", tau, tau))
    if reduction == "mean":
        s = s / math.sqrt(n_total)
    return s


def sample_ce_classes(
    logits: Tensor, subsampling: Optional[Sequence[int]], mc_samples: int
) -> Tensor:
    """Class ids ``y[m,n] ~ Cat(softmax(logits_n))`` -- the one random draw of the
    MC factor; ``multinomial(replacement=True)`` on torch's global generator."""
    probs = _sub(F.softmax(logits, dim=1), subsampling)
    return torch.multinomial(probs, mc_samples, replacement=True).t().contiguous()


def sqrt_hessian_ce_sampled(
    logits: Tensor,
    subsampling: Optional[Sequence[int]],
    class_ids: Tensor,
    reduction: str = "mean",
) -> Tensor:
    """``S[m,n,c] = (p_nc - 1[y_mn = c]) / sqrt(M N)``."""
    n_total = logits.shape[0]
    probs = _sub(F.softmax(logits, dim=1), subsampling)
    m = class_ids.shape[0]
    onehot = F.one_hot(class_ids, probs.shape[1]).to(probs.dtype)  # [M, N, C]
    s = (probs[None] - onehot) / math.sqrt(m)
    if reduction == "mean":
        s = s / math.sqrt(n_total)
    return s


def sqrt_hessian_mse(
    out: Tensor, subsampling: Optional[Sequence[int]], reduction: str = "mean"
) -> Tensor:
    """``sqrt(2) I_C`` per sample, divided by ``sqrt(N C)`` for ``mean``."""
    if out.dim() != 2:
        raise ValueError("MSE factor only supports 2d inputs")
    n_sub = out.shape[0] if subsampling is None else len(subsampling)
    c = out.shape[1]
    s = math.sqrt(2.0) * torch.eye(c, dtype=out.dtype, device=out.device)
    s = s[:, None, :].expand(c, n_sub, c).clone()
    if reduction == "mean":
        s = s / math.sqrt(out.numel())
    return s


def sqrt_hessian_mse_sampled(
    out: Tensor,
    subsampling: Optional[Sequence[int]],
    normal: Tensor,
    reduction: str = "mean",
) -> Tensor:
    """``normal[M,N,C] * sqrt(2/M)`` (``/ sqrt(N C)`` for ``mean``)."""
    m = normal.shape[0]
    s = normal * (math.sqrt(2.0) / math.sqrt(m))
    if reduction == "mean":
        s = s / math.sqrt(out.numel())
    return s


def loss_sqrt_hessian(
    loss_fn: nn.Module,
    model_out: Tensor,
    subsampling: Optional[Sequence[int]],
    mc_samples: int = 0,
    mc_state: Optional[Tensor] = None,
) -> Tensor:
    """Dispatch on loss type / exact-vs-MC (``secondorder/vivit/__init__.py:80-86,122-128``).

    ``mc_state`` carries the random draw (class ids for CE, normal samples for
    MSE) so the CUDA path and the oracle can be compared on identical samples.
    """
    red = loss_fn.reduction
    if isinstance(loss_fn, nn.CrossEntropyLoss):
        if mc_samples == 0:
            return sqrt_hessian_ce(model_out, subsampling, red)
        if mc_state is None:
            mc_state = sample_ce_classes(model_out, subsampling, mc_samples)
        return sqrt_hessian_ce_sampled(model_out, subsampling, mc_state, red)
    if isinstance(loss_fn, nn.MSELoss):
        if mc_samples == 0:
            return sqrt_hessian_mse(model_out, subsampling, red)
        if mc_state is None:
            n_sub = model_out.shape[0] if subsampling is None else len(subsampling)
            mc_state = torch.randn(
                mc_samples, n_sub, model_out.shape[1], dtype=model_out.dtype
            )
        return sqrt_hessian_mse_sampled(model_out, subsampling, mc_state, red)
    raise NotImplementedError(f"loss {type(loss_fn)}")


# --------------------------------------------------------------------------
# a4: transposed-Jacobian products of layers  [external]
# --------------------------------------------------------------------------


def jac_t_mat_prod(
    module: nn.Module, inp: Tensor, out: Tensor, mat: Tensor
) -> Tensor:
    """``[V, N, *out] -> [V, N, *in]``; ``inp``/``out`` already sub-sampled."""
    v, n = mat.shape[:2]
    if isinstance(module, nn.Linear):
        return einsum("vn...o,oi->vn...i", mat, module.weight)
    if isinstance(module, nn.Conv2d):
        flat = mat.reshape(v * n, *mat.shape[2:])
        h, w = inp.shape[2:]
        # output_padding resolves the stride ambiguity of the transposed conv
        opad = []
        for d, size in enumerate((h, w)):
            base = (
                (mat.shape[3 + d] - 1) * module.stride[d]
                - 2 * module.padding[d]
                + module.dilation[d] * (module.kernel_size[d] - 1)
                + 1
            )
            opad.append(max(0, min(size - base, max(module.stride[d], module.dilation[d]) - 1)))
        res = F.conv_transpose2d(
            flat,
            module.weight,
            None,
            stride=module.stride,
            padding=module.padding,
            output_padding=tuple(opad),
            dilation=module.dilation,
            groups=module.groups,
        )
        if res.shape[2] < h or res.shape[3] < w:  # trailing inputs no window touches
            res = F.pad(res, (0, w - res.shape[3], 0, h - res.shape[2]))
        res = res[:, :, :h, :w]
        return res.reshape(v, n, *inp.shape[1:])
    if isinstance(module, nn.ReLU):
        return mat * (inp > 0).to(mat.dtype)[None]
    if isinstance(module, nn.Sigmoid):
        return mat * (out * (1.0 - out))[None]
    if isinstance(module, nn.Tanh):
        return mat * (1.0 - out**2)[None]
    if isinstance(module, (nn.LeakyReLU, nn.ELU, nn.SELU, nn.LogSigmoid)):
        # [external] ElementwiseDerivatives: mat * f'(input); f' by autograd of the module itself
        x = inp.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            (d,) = torch.autograd.grad(module(x).sum(), x)
        return mat * d[None]
    if isinstance(module, (nn.Flatten,)):
        return mat.reshape(v, n, *inp.shape[1:])
    if isinstance(module, nn.Identity):
        return mat
    if type(module).__name__ == "ScaleModule" and isinstance(getattr(module, "weight", None), float):
        return mat * module.weight  # [external] ScaleModuleDerivatives
    if type(module).__name__ == "Pad" and hasattr(module, "pad"):
        # [external] PadDerivatives: constant padding (any fill value) has a zero Jacobian on the border
        if module.mode != "constant":
            raise NotImplementedError("Pad: mode='constant' only")
        res = mat
        for i in range(len(module.pad) // 2):
            axis = mat.dim() - 1 - i
            res = res.narrow(axis, module.pad[2 * i], mat.shape[axis] - module.pad[2 * i] - module.pad[2 * i + 1])
        return res
    if type(module).__name__ == "Slicing" and hasattr(module, "slice_info"):
        # [external] SlicingDerivatives: scatter the rows back, zeros where the slice dropped entries
        res = torch.zeros(v, n, *inp.shape[1:], dtype=mat.dtype, device=mat.device)
        res[(slice(None), slice(None)) + tuple(module.slice_info[1:])] = mat
        return res
    if isinstance(module, nn.Dropout):
        if not module.training or module.p == 0.0:
            return mat
        mask = (out != 0).to(mat.dtype) / (1.0 - module.p)
        return mat * mask[None]
    if isinstance(module, nn.MaxPool2d):
        _, idx = F.max_pool2d(
            inp,
            module.kernel_size,
            module.stride,
            module.padding,
            module.dilation,
            module.ceil_mode,
            return_indices=True,
        )
        ch, h, w = inp.shape[1:]
        flat = mat.reshape(v, n, ch, -1)
        res = torch.zeros(v, n, ch, h * w, dtype=mat.dtype, device=mat.device)
        res.scatter_add_(3, idx.reshape(1, n, ch, -1).expand(v, -1, -1, -1), flat)
        return res.reshape(v, n, ch, h, w)
    if isinstance(module, nn.AvgPool2d):
        # linear layer: use autograd of the layer itself on the V*N "virtual batch"
        x = torch.zeros(v * n, *inp.shape[1:], dtype=mat.dtype, requires_grad=True)
        with torch.enable_grad():
            y = module(x)
            (res,) = torch.autograd.grad(y, x, mat.reshape(v * n, *mat.shape[2:]))
        return res.reshape(v, n, *inp.shape[1:])
    if isinstance(module, nn.ZeroPad2d):
        left, right, top, bottom = module.padding
        h, w = mat.shape[-2:]
        return mat[..., top : h - bottom, left : w - right]
    if isinstance(module, (nn.Conv1d, nn.MaxPool1d, nn.AvgPool1d, nn.Conv3d, *_CONV_TRANSPOSE)):
        # [external] Conv1D/3DDerivatives / ConvTransposeNDDerivatives / MaxPool1DDerivatives /
        # AvgPool1DDerivatives: transposed Jacobian of the layer at its forward input, by autograd of the layer
        # itself on the V*N "virtual batch"
        x = inp.repeat(v, *([1] * (inp.dim() - 1))).detach().clone().requires_grad_(True)
        with torch.enable_grad():
            y = module(x)
            (res,) = torch.autograd.grad(y, x, mat.reshape(v * n, *mat.shape[2:]))
        return res.reshape(v, n, *inp.shape[1:])
    if isinstance(module, _BATCHNORM):
        # [external] BatchNormNdDerivatives._jac_t_mat_prod in evaluation mode: the layer is an affine map
        # per channel, its Jacobian is diag(weight / sqrt(running_var + eps))
        _bn_require_eval(module)
        shape = (1, 1, -1) + (1,) * (mat.dim() - 3)
        return mat * (module.weight / (module.running_var + module.eps).sqrt()).reshape(shape)
    raise NotImplementedError(f"jac_t_mat_prod for {type(module)}")


# --------------------------------------------------------------------------
# a6: per-sample parameter Jacobian products  [external param_mjp]
# --------------------------------------------------------------------------


def param_mjp(
    module: nn.Module, name: str, inp: Tensor, mat: Tensor
) -> Tensor:
    """``[V, N, *out] -> [V, N, *param.shape]`` (``sum_batch=False``),
    as called from ``vivit/extensions/secondorder/vivit/base.py:84-92``."""
    if isinstance(module, nn.Linear):
        if name == "weight":
            return einsum("vn...o,n...i->vnoi", mat, inp)
        extra = tuple(range(2, mat.dim() - 1))
        return mat.sum(extra) if extra else mat
    if isinstance(module, nn.Conv2d):
        if name == "bias":
            return mat.sum((3, 4))
        if module.groups != 1:
            raise NotImplementedError("grouped convolution")
        v, n, co = mat.shape[:3]
        cols = F.unfold(
            inp,
            module.kernel_size,
            dilation=module.dilation,
            padding=module.padding,
            stride=module.stride,
        )  # [N, J, X]
        vt = einsum("vnox,njx->vnoj", mat.reshape(v, n, co, -1), cols)
        return vt.reshape(v, n, *module.weight.shape)
    if isinstance(module, (nn.Conv1d, nn.Conv3d, *_CONV_TRANSPOSE)):
        # [external] Conv1D/3DDerivatives.param_mjp, ConvTransposeNDDerivatives.param_mjp: the layer is linear in
        # its parameters, so the product is the gradient of <mat[v, n], layer(x_n)> -- one autograd call per
        # (v, n), fine at oracle sizes
        if name == "bias":
            return mat.sum(tuple(range(3, mat.dim())))
        v, n = mat.shape[:2]
        rows = []
        for vi in range(v):
            for ni in range(n):
                w = module.weight.detach().clone().requires_grad_(True)
                with torch.enable_grad():
                    y = torch.func.functional_call(module, {"weight": w}, (inp[ni : ni + 1],))
                    (g,) = torch.autograd.grad(y, w, mat[vi, ni][None])
                rows.append(g)
        return torch.stack(rows).reshape(v, n, *module.weight.shape)
    if isinstance(module, _BATCHNORM):
        # [external] BatchNormNdDerivatives._weight_jac_t_mat_prod / _bias_jac_t_mat_prod (evaluation mode),
        # wired by ``vivit/extensions/secondorder/vivit/batchnormnd.py:8-13`` (params=["bias", "weight"])
        _bn_require_eval(module)
        spatial = tuple(range(3, mat.dim()))
        if name == "bias":
            return mat.sum(spatial) if spatial else mat
        x_hat = F.batch_norm(inp, module.running_mean, module.running_var, None, None, False, 0.0, module.eps)
        prod = mat * x_hat[None]
        return prod.sum(spatial) if spatial else prod
    raise NotImplementedError(f"param_mjp for {type(module)}")


_BATCHNORM = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)
_CONV_TRANSPOSE = (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)


def _bn_require_eval(module: nn.Module) -> None:
    if module.training or not module.track_running_stats or module.weight is None:
        raise NotImplementedError("BatchNorm is supported in evaluation mode (affine, running statistics) only")


def _has_additional_dims(inp: Tensor) -> bool:
    """``LinearDerivatives._get_additional_dims`` [external] (``linear.py:38``)."""
    return inp.dim() > 2


# --------------------------------------------------------------------------
# a7/a8: contraction primitives (vivit/utils/gram.py, vivit/utils/ggn.py)
# --------------------------------------------------------------------------


def partial_contract(t: Tensor, o: Tensor, start_dims: Tuple[int, int]) -> Tensor:
    """``vivit/utils/gram.py:206-232``: contract all trailing dims."""
    d1, d2 = start_dims
    a = t.flatten(start_dim=d1) if t.dim() > d1 else t[..., None]
    b = o.flatten(start_dim=d2) if o.dim() > d2 else o[..., None]
    res = a.reshape(-1, a.shape[-1]) @ b.reshape(-1, b.shape[-1]).t()
    return res.reshape(*t.shape[:d1], *o.shape[:d2])


def pairwise_dot(t: Tensor, start_dim: int) -> Tensor:
    """``vivit/utils/gram.py:9-35`` with ``flatten=False``."""
    return partial_contract(t, t, (start_dim, start_dim))


def reshape_as_square(t: Tensor) -> Tensor:
    """``vivit/utils/gram.py:58-69``."""
    dim = int(math.sqrt(t.numel()))
    return t.reshape(dim, dim)


def Vmp(V_t: Tensor, mat: Tensor, start_dim: int = 2) -> Tensor:
    """``V @ mat``: ``[F, C, N] x [C, N, *p] -> [F, *p]`` (``vivit/utils/ggn.py:94-115``)."""
    lead = V_t.shape[:start_dim].numel()
    res = mat.reshape(mat.shape[0], lead) @ V_t.reshape(lead, -1)
    return res.reshape(mat.shape[0], *V_t.shape[start_dim:])


def mVp(V_t: Tensor, mat: Tensor, start_dim: int = 2) -> Tensor:
    """``V^T @ mat``: ``[F, *p] x [C, N, *p] -> [F, C, N]`` (``vivit/utils/gram.py:182-203``)."""
    lead = V_t.shape[:start_dim].numel()
    res = mat.reshape(mat.shape[0], -1) @ V_t.reshape(lead, -1).t()
    return res.reshape(mat.shape[0], *V_t.shape[:start_dim])


# --------------------------------------------------------------------------
# a5/a6: the ViViTGGN{Exact,MC} savefield closures
# --------------------------------------------------------------------------


def _linear_weight_closures(s: Tensor, z: Tensor) -> Dict[str, Callable]:
    """Structured path of ``nn.Linear.weight`` with 2d input
    (``vivit/extensions/secondorder/vivit/linear.py:41-81``)."""

    def V_mat_prod(mat: Tensor) -> Tensor:  # linear.py:44-53
        return einsum("cno,vcn,ni->voi", s, mat, z)

    def V_t_mat_prod(mat: Tensor) -> Tensor:  # linear.py:55-64
        return einsum("cno,voi,ni->vcn", s, mat, z)

    def gram_mat() -> Tensor:  # linear.py:66-75
        s2 = pairwise_dot(s, 2)
        z2 = pairwise_dot(z, 1)
        return einsum("nm,cndm->cndm", z2, s2)

    return {"V_mat_prod": V_mat_prod, "V_t_mat_prod": V_t_mat_prod, "gram_mat": gram_mat}


def _dense_closures(V_t: Tensor) -> Dict[str, Callable]:
    """Generic materialised path (``vivit/extensions/secondorder/vivit/base.py:94-130``)."""
    return {
        "V_mat_prod": lambda mat: Vmp(V_t, mat, 2),
        "V_t_mat_prod": lambda mat: mVp(V_t, mat, 2),
        "gram_mat": lambda: pairwise_dot(V_t, 2),
        "_V_t": V_t,
    }


def _param_items(module: nn.Module):
    # BackPACK applies the param functions in the order given at construction:
    # ["bias", "weight"] (linear.py:24, convnd.py:14).  Only matters for fp32
    # summation order.
    for name in ("bias", "weight"):
        p = getattr(module, name, None)
        if isinstance(p, nn.Parameter) and p.requires_grad:
            yield name, p


class BackwardResult:
    """Per-parameter quantities gathered in one restated backward sweep."""

    def __init__(self):
        self.order: List[nn.Parameter] = []  # order in which hooks would fire
        self.vivit: Dict[int, Dict[str, Callable]] = {}
        self.sqrt_ggn: Dict[int, Tensor] = {}
        self.grad_batch: Dict[int, Tensor] = {}
        self.batch_size: int = 0


def backward_sweep(
    model: nn.Module,
    loss_fn: nn.Module,
    x: Tensor,
    y: Tensor,
    *,
    subsampling_ggn: Optional[Sequence[int]] = None,
    mc_samples: int = 0,
    mc_state: Optional[Tensor] = None,
    want_vivit: bool = False,
    want_sqrt_ggn: bool = False,
    want_grad_batch: bool = False,
    subsampling_grad: Optional[Sequence[int]] = None,
) -> BackwardResult:
    """One pass of what BackPACK does during ``loss.backward()`` with the
    ViViTGGN / SqrtGGN / BatchGrad extensions active."""
    res = BackwardResult()
    res.batch_size = x.shape[0]

    if want_grad_batch:
        records, h = forward_capture(model, x.detach(), retain_grad=True)
        loss = loss_fn(h, y)
        loss.backward()
        out = h.detach()
        grads = _collect_grads(records)
        for p in model.parameters():
            p.grad = None
    else:
        with torch.no_grad():
            records, out = forward_capture(model, x)
        grads = {}

    need_s = want_vivit or want_sqrt_ggn
    s = None
    if need_s:
        with torch.no_grad():
            s = loss_sqrt_hessian(loss_fn, out, subsampling_ggn, mc_samples, mc_state)

    def visit(module, inp, outp, s):
        inp, outp = inp.detach(), outp.detach()
        inp_s = _sub(inp, subsampling_ggn)
        for name, p in _param_items(module):
            res.order.append(p)
            if need_s:
                structured = (
                    isinstance(module, nn.Linear)
                    and name == "weight"
                    and not _has_additional_dims(inp)
                )
                if want_vivit:
                    if structured:
                        res.vivit[id(p)] = _linear_weight_closures(s, inp_s)
                    else:
                        res.vivit[id(p)] = _dense_closures(
                            param_mjp(module, name, inp_s, s)
                        )
                if want_sqrt_ggn:
                    res.sqrt_ggn[id(p)] = param_mjp(module, name, inp_s, s)
            if want_grad_batch:
                g = _sub(grads[id(module)], subsampling_grad)[None]
                res.grad_batch[id(p)] = param_mjp(
                    module, name, _sub(inp, subsampling_grad), g
                )[0]

    def sweep(records, s, first):
        """Reverse pass over one chain; returns the factor at the chain's input.  The sum over the
        branches of a ``parallel`` record is ``accumulate_backpropagated_quantities``
        (``secondorder/vivit/__init__.py:130-133``)."""
        for rec in reversed(records):
            is_first = first and rec is records[0]
            if rec[0] == "parallel":
                parts = [sweep(branch, s, False) for branch in rec[1]]
                if need_s:
                    s = parts[0]
                    for part in parts[1:]:
                        s = s + part
                continue
            _, module, inp, outp = rec
            visit(module, inp, outp, s)
            if need_s and not is_first:
                s = jac_t_mat_prod(
                    module, _sub(inp.detach(), subsampling_ggn), _sub(outp.detach(), subsampling_ggn), s
                )
        return s

    with torch.no_grad():
        sweep(records, s, True)
    return res


def _collect_grads(records) -> Dict[int, Tensor]:
    """``d loss / d output`` of every leaf with parameters, keyed by ``id(module)``."""
    grads: Dict[int, Tensor] = {}
    for rec in records:
        if rec[0] == "parallel":
            for branch in rec[1]:
                grads.update(_collect_grads(branch))
        elif rec[3].grad is not None and any(True for _ in _param_items(rec[1])):
            grads[id(rec[1])] = rec[3].grad.detach()
    return grads


# --------------------------------------------------------------------------
# a9-a12: the four Computations
# --------------------------------------------------------------------------


def eigh_psd(mat: Tensor, eigenvectors: bool = True):
    """Stand-in for ``Tensor.symeig(eigenvectors, upper=True)``."""
    if eigenvectors:
        return torch.linalg.eigh(mat, UPLO="U")
    return torch.linalg.eigvalsh(mat, UPLO="U"), None


def normalize(tensors: List[Tensor]) -> List[Tensor]:
    """``vivit/linalg/utils.py:67-76``."""
    inv = 1.0 / sum((t**2).flatten(1).sum(1) for t in tensors).sqrt()
    return [einsum("i,i...->i...", inv, t) for t in tensors]


def remove_zero_evals(evals: Tensor, evecs: Tensor, atol=1e-7, rtol=1e-5):
    """``vivit/utils/eig.py:111-134``."""
    keep = torch.isclose(evals, torch.zeros_like(evals), rtol=rtol, atol=atol).logical_not()
    return evals[keep], (evecs[:, keep] if evecs.numel() else evecs)


def eigvalsh(
    model, loss_fn, x, y, param_groups, subsampling=None, mc_samples=0, mc_state=None
) -> List[Tensor]:
    """``EigvalshComputation`` (``vivit/linalg/eigvalsh.py:145-158,170-183,201-225``)."""
    sweep = backward_sweep(
        model, loss_fn, x, y, subsampling_ggn=subsampling, mc_samples=mc_samples,
        mc_state=mc_state, want_vivit=True,
    )
    n = sweep.batch_size
    out = []
    for group in param_groups:
        ids = {id(p) for p in group["params"]}
        gram = None
        for p in sweep.order:  # eager, in hook order (eigvalsh.py:154, :183)
            if id(p) in ids:
                g = sweep.vivit[id(p)]["gram_mat"]()
                gram = g if gram is None else gram + g
        gram = reshape_as_square(gram)
        if subsampling is not None:
            gram = gram * (n / len(subsampling))  # eigvalsh.py:218-219
        evals, _ = eigh_psd(gram, eigenvectors=False)
        out.append(evals)
    return out


def eigh(
    model, loss_fn, x, y, param_groups, subsampling=None, mc_samples=0, mc_state=None
) -> List[Tuple[Tensor, List[Tensor]]]:
    """``EighComputation`` group hook (``vivit/linalg/eigh.py:232-275``)."""
    sweep = backward_sweep(
        model, loss_fn, x, y, subsampling_ggn=subsampling, mc_samples=mc_samples,
        mc_state=mc_state, want_vivit=True,
    )
    n = sweep.batch_size
    out = []
    for group in param_groups:
        gram = 0.0
        for p in group["params"]:  # eigh.py:241-242
            gram = gram + sweep.vivit[id(p)]["gram_mat"]()
        if subsampling is not None:
            gram = gram * (n / len(subsampling))  # eigh.py:245-246
        evals, evecs = eigh_psd(reshape_as_square(gram))
        keep = group["criterion"](evals)
        evals, evecs = evals[keep], evecs[:, keep]
        evecs = evecs.transpose(0, 1).reshape(-1, *gram.shape[:2])  # eigh.py:265
        group_evecs = [sweep.vivit[id(p)]["V_mat_prod"](evecs) for p in group["params"]]
        out.append((evals, normalize(group_evecs)))
    return out


def _gram_space_quantities(sweep: BackwardResult, group):
    """Per-param dot products and their accumulation
    (``vivit/optim/directional_derivatives.py:238-247,348-351``)."""
    ids = {id(p) for p in group["params"]}
    vtv = vtg = None
    for p in sweep.order:
        if id(p) not in ids:
            continue
        V, g = sweep.sqrt_ggn[id(p)], sweep.grad_batch[id(p)]
        a = partial_contract(V, V, (2, 2))
        b = partial_contract(V, g, (2, 1))
        vtv = a if vtv is None else vtv + a
        vtg = b if vtg is None else vtg + b
    return vtv, vtg


def _directions(sweep, group, vtv, vtg):
    """Shared first half of both optim group hooks
    (``directional_derivatives.py:281-325`` == ``directional_damped_newton.py:304-351``)."""
    N = sweep.batch_size
    N_ggn = vtv.shape[1]
    corr = math.sqrt(N / N_ggn)
    gram = corr**2 * vtv
    evals, evecs = eigh_psd(reshape_as_square(gram))
    keep = group["criterion"](evals)
    evals, evecs = evals[keep], evecs[:, keep]
    V_t_g_n = corr * N * vtg.flatten(0, 1)
    gammas = einsum("in,id->nd", V_t_g_n, evecs) / evals.sqrt()
    V_n_T_V_e_d = math.sqrt(N_ggn) * einsum("cni,id->cnd", gram.flatten(2), evecs)
    lambdas = (V_n_T_V_e_d**2).sum(0) / evals
    return corr, gram, evals, evecs, gammas, lambdas


def directional_derivatives(
    model, loss_fn, x, y, param_groups, subsampling_grad=None, subsampling_ggn=None,
    mc_samples_ggn=0, mc_state=None,
) -> List[Tuple[Tensor, Tensor]]:
    """``DirectionalDerivativesComputation`` -> ``(gammas[N_grad,K], lambdas[N_ggn,K])``."""
    sweep = backward_sweep(
        model, loss_fn, x, y, subsampling_ggn=subsampling_ggn, mc_samples=mc_samples_ggn,
        mc_state=mc_state, want_sqrt_ggn=True, want_grad_batch=True,
        subsampling_grad=subsampling_grad,
    )
    out = []
    for group in param_groups:
        vtv, vtg = _gram_space_quantities(sweep, group)
        _, _, _, _, gammas, lambdas = _directions(sweep, group, vtv, vtg)
        out.append((gammas, lambdas))
    return out


def directional_damped_newton(
    model, loss_fn, x, y, param_groups, subsampling_grad=None, subsampling_ggn=None,
    mc_samples_ggn=0, mc_state=None,
) -> List[List[Tensor]]:
    """``DirectionalDampedNewtonComputation`` group hook
    (``vivit/optim/directional_damped_newton.py:304-379``)."""
    sweep = backward_sweep(
        model, loss_fn, x, y, subsampling_ggn=subsampling_ggn, mc_samples=mc_samples_ggn,
        mc_state=mc_state, want_sqrt_ggn=True, want_grad_batch=True,
        subsampling_grad=subsampling_grad,
    )
    out = []
    for group in param_groups:
        vtv, vtg = _gram_space_quantities(sweep, group)
        corr, gram, evals, evecs, gammas, lambdas = _directions(sweep, group, vtv, vtg)
        C, N_ggn = gram.shape[:2]
        deltas = group["damping"](evals, evecs, gammas, lambdas)
        coeff = -gammas.mean(0) / (lambdas.mean(0) + deltas) / evals.sqrt()  # :354-359
        v = einsum("id,d->i", evecs, coeff) * corr  # :362-366
        v = v.reshape(C, N_ggn)
        out.append(
            [einsum("cn,cn...->...", v, sweep.sqrt_ggn[id(p)]) for p in group["params"]]
        )
    return out


# --------------------------------------------------------------------------
# f2: Gram hooks on materialised quantities (vivit/extensions/hooks.py)
# --------------------------------------------------------------------------


def centered_batch_grad(model: nn.Module, loss_fn: nn.Module, x: Tensor, y: Tensor) -> List[Tensor]:
    """``CenteredBatchGrad.param_hook``
    (``vivit/extensions/firstorder/batch_grad/gram_batch_grad.py:25-38``): per parameter (in
    ``model.parameters()`` order) ``grad_batch - grad_batch.mean(0)``."""
    sweep = backward_sweep(model, loss_fn, x, y, want_grad_batch=True)
    out = []
    for p in model.parameters():
        if p.requires_grad:
            gb = sweep.grad_batch[id(p)]
            out.append(gb - gb.mean(0))
    return out


def gram_batch_grad(
    model: nn.Module, loss_fn: nn.Module, x: Tensor, y: Tensor, center: bool = False
) -> Tuple[Tensor, Dict[int, Tensor]]:
    """``_GramBatchGradBase.param_hook`` / ``get_result`` (``gram_batch_grad.py:73-123``):
    the ``[N, N]`` sum over parameters of ``pairwise_dot(grad_batch, start_dim=1)``, after
    ``grad_batch -= grad_batch.mean(0)`` if ``center``; also the per-parameter (layer-wise)
    matrices keyed by ``id(param)``, in the order the hook would visit the parameters."""
    sweep = backward_sweep(model, loss_fn, x, y, want_grad_batch=True)
    total, layers = None, {}
    for p in sweep.order:
        gb = sweep.grad_batch[id(p)]
        if center:
            gb = gb - gb.mean(0)
        gram = reshape_as_square(pairwise_dot(gb, 1))
        layers[id(p)] = gram
        total = gram.clone() if total is None else total + gram
    return total, layers


def gram_sqrt_ggn(
    model: nn.Module,
    loss_fn: nn.Module,
    x: Tensor,
    y: Tensor,
    mc_samples: int = 0,
    mc_state: Optional[Tensor] = None,
    subsampling: Optional[Sequence[int]] = None,
) -> Tuple[Tensor, Dict[int, Tensor]]:
    """``GramSqrtGGN.param_hook`` / ``get_result``
    (``vivit/extensions/secondorder/sqrt_ggn/gram_sqrt_ggn.py:41-74``): the ``[CN, CN]`` sum over
    parameters of ``pairwise_dot(sqrt_ggn, start_dim=2)`` flattened to a square, plus the
    layer-wise matrices."""
    sweep = backward_sweep(
        model, loss_fn, x, y, subsampling_ggn=subsampling, mc_samples=mc_samples,
        mc_state=mc_state, want_sqrt_ggn=True,
    )
    total, layers = None, {}
    for p in sweep.order:
        gram = reshape_as_square(pairwise_dot(sweep.sqrt_ggn[id(p)], 2))
        layers[id(p)] = gram
        total = gram.clone() if total is None else total + gram
    return total, layers
