"""Curvature matrices of a network's loss as matrix-free linear operators over a data set."""

from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch
from scipy.sparse.linalg import LinearOperator
from torch import Tensor, nn


def _filled(grads: Sequence[Tensor], like: Sequence[Tensor]) -> List[Tensor]:
    """Replace the ``None`` of unused parameters by zeros."""
    return [torch.zeros_like(p) if g is None else g for g, p in zip(grads, like)]


def hessian_vector_product(loss: Tensor, params: Sequence[Tensor], v: Sequence[Tensor]) -> List[Tensor]:
    """``(d^2 loss / d params^2) v`` by differentiating ``<grad, v>`` (one double-backward pass)."""
    grads = torch.autograd.grad(loss, params, create_graph=True, allow_unused=True)
    pairs = [(g, x) for g, x in zip(grads, v) if g is not None and g.requires_grad]
    if not pairs:
        return [torch.zeros_like(p) for p in params]
    dot = sum((g * x).sum() for g, x in pairs)
    return _filled(torch.autograd.grad(dot, params, allow_unused=True), params)


def ggn_vector_product(loss: Tensor, output: Tensor, params: Sequence[Tensor], v: Sequence[Tensor]) -> List[Tensor]:
    """``J^T (d^2 loss / d output^2) J v`` with ``J = d output / d params``.

    ``J v`` comes from the transposed product twice (``u -> J^T u`` is linear, so differentiating
    ``<J^T u, v>`` w.r.t. ``u`` gives ``J v``), the middle factor is a Hessian-vector product of the loss
    w.r.t. the model output, and ``J^T`` is one more backward pass.
    """
    u = torch.zeros_like(output, requires_grad=True)
    jt_u = torch.autograd.grad(output, params, grad_outputs=u, create_graph=True, allow_unused=True)
    dot = sum((g * x).sum() for g, x in zip(jt_u, v) if g is not None)
    (jv,) = torch.autograd.grad(dot, u)
    (g_out,) = torch.autograd.grad(loss, output, create_graph=True)
    if g_out.requires_grad:
        (h_jv,) = torch.autograd.grad(g_out, output, grad_outputs=jv, retain_graph=True)
    else:  # loss linear in the output
        h_jv = torch.zeros_like(output)
    return _filled(torch.autograd.grad(output, params, grad_outputs=h_jv.detach(), allow_unused=True), params)


class _CurvatureOperator(LinearOperator):
    """``dim x dim`` curvature matrix of ``sum/mean_n loss(model(x_n), y_n)`` over all mini-batches of ``data``
    (``vivit/hessianfree/__init__.py:21-273``).  Usable with SciPy (``A @ v``, ``eigsh(A)``) and, without host
    round trips, through ``matvec_torch``."""

    def __init__(
        self,
        model: nn.Module,
        loss_func: nn.Module,
        data: Iterable[Tuple[Tensor, Tensor]],
        device: torch.device,
        dtype=np.float32,
        progressbar: bool = False,
        check_deterministic: bool = True,
    ):
        self._params = [p for p in model.parameters() if p.requires_grad]
        dim = sum(p.numel() for p in self._params)
        super().__init__(shape=(dim, dim), dtype=dtype)
        self._model, self._loss_func, self._data = model, loss_func, data
        self._progressbar = progressbar
        self.to_device(torch.device(device))
        self._N_data = sum(X.shape[0] for X, _ in self._batches())
        if check_deterministic:
            home = self._device
            self.to_device(torch.device("cpu"))  # as the reference: GPU reductions may be non-deterministic
            try:
                self._check_deterministic()
            finally:
                self.to_device(home)

    # -- plumbing ---------------------------------------------------------------------------------
    def to_device(self, device: torch.device) -> None:
        self._device = device
        self._model = self._model.to(device)
        self._loss_func = self._loss_func.to(device)

    def _batches(self):
        it = iter(self._data)
        if self._progressbar:
            from tqdm import tqdm

            it = tqdm(it, desc="matvec")
        for X, y in it:
            yield X.to(self._device), y.to(self._device)

    def _weight(self, X: Tensor) -> float:
        """Factor that turns the per-batch reduction into the reduction over the whole data set."""
        reduction = getattr(self._loss_func, "reduction", None)
        if reduction is None:
            raise ValueError("Loss must have a 'reduction' attribute.")
        if reduction == "sum":
            return 1.0
        if reduction == "mean":
            return X.shape[0] / self._N_data
        raise ValueError("Loss must have reduction 'mean' or 'sum'.")

    def _split(self, flat: Tensor) -> List[Tensor]:
        out, start = [], 0
        for p in self._params:
            out.append(flat[start : start + p.numel()].reshape(p.shape))
            start += p.numel()
        return out

    # -- products ---------------------------------------------------------------------------------
    def _matvec_batch(self, X: Tensor, y: Tensor, x_list: List[Tensor]) -> List[Tensor]:
        raise NotImplementedError

    def matvec_torch(self, v: Tensor) -> Tensor:
        """``A v`` for a flat tensor on the operator's device; the result stays there."""
        like = self._params[0]
        x_list = self._split(v.to(device=self._device, dtype=like.dtype))
        acc = [torch.zeros_like(p) for p in self._params]
        for X, y in self._batches():
            w = self._weight(X)
            for a, cur in zip(acc, self._matvec_batch(X, y, x_list)):
                a.add_(cur, alpha=w)
        return torch.cat([a.reshape(-1) for a in acc])

    def _matvec(self, x: np.ndarray) -> np.ndarray:
        v = torch.from_numpy(np.ascontiguousarray(x).reshape(-1))
        return self.matvec_torch(v).cpu().numpy().astype(self.dtype, copy=False)

    def gradient_and_loss(self) -> Tuple[List[Tensor], Tensor]:
        """Gradient (parameter-list format) and loss over the data set (``__init__.py:227-246``)."""
        total_loss = torch.zeros(1, device=self._device)
        total_grad = [torch.zeros_like(p) for p in self._params]
        for X, y in self._batches():
            loss = self._loss_func(self._model(X), y)
            w = self._weight(X)
            for g, cur in zip(total_grad, _filled(torch.autograd.grad(loss, self._params, allow_unused=True), self._params)):
                g.add_(cur, alpha=w)
            total_loss.add_(loss.detach().to(total_loss.dtype), alpha=w)
        return total_grad, total_loss

    # -- safeguard ----------------------------------------------------------------------------------
    def _check_deterministic(self) -> None:
        """Two evaluations of loss, gradient and a matrix-vector product must agree (``__init__.py:89-133``):
        catches data augmentation, ``drop_last`` shuffling, Dropout / BatchNorm in training mode."""
        rtol, atol = 5e-5, 1e-6

        def flat(ts):
            return torch.cat([t.reshape(-1) for t in ts]).cpu().numpy()

        (g1, l1), (g2, l2) = self.gradient_and_loss(), self.gradient_and_loss()
        if not np.allclose(l1.cpu().numpy(), l2.cpu().numpy(), rtol=rtol, atol=atol):
            raise RuntimeError("Check for deterministic loss failed.")
        if not np.allclose(flat(g1), flat(g2), rtol=rtol, atol=atol):
            raise RuntimeError("Check for deterministic gradient failed.")
        v = np.random.rand(self.shape[0]).astype(self.dtype)
        if not np.allclose(self @ v, self @ v, rtol=rtol, atol=atol):
            raise RuntimeError("Check for deterministic matvec failed.")


class HessianLinearOperator(_CurvatureOperator):
    """Hessian of the loss over the data set (``vivit/hessianfree/__init__.py:276-295``)."""

    def _matvec_batch(self, X, y, x_list):
        loss = self._loss_func(self._model(X), y)
        return hessian_vector_product(loss, self._params, x_list)


class GGNLinearOperator(_CurvatureOperator):
    """Generalized Gauss-Newton matrix over the data set (``vivit/hessianfree/__init__.py:298-318``)."""

    def _matvec_batch(self, X, y, x_list):
        output = self._model(X)
        loss = self._loss_func(output, y)
        return ggn_vector_product(loss, output, self._params, x_list)
