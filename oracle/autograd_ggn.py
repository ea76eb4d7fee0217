"""Brute-force ``torch.autograd`` ground truth for the GGN hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Restates what the reference's tests use as the pinning side of every hot-path
comparison (``test/implementation/autograd.py``, ``linalg_autograd.py``,
``optim_autograd.py``): the full GGN ``J^T H J`` w.r.t. the parameters, the
per-sample GGNs, per-sample gradients, and from those the eigenpairs, the
directional derivatives and the damped Newton step.  Only feasible for the tiny
problems of ``test/settings.py`` (D up to a few hundred).

Where the reference builds the GGN row by row from GGN-vector products
(``autograd.py:74-93``), this builds the output Jacobian ``J`` and the loss
Hessian ``H`` explicitly and multiplies -- the same matrix.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from oracle.reference_path import eigh_psd


def _flat_params(params: Sequence[Tensor]) -> int:
    return sum(p.numel() for p in params)


def output_jacobian(model: nn.Module, x: Tensor, params: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """``J[(n, c...), d] = d out[n, c...] / d theta_d`` and the output itself."""
    out = model(x)
    flat = out.reshape(-1)
    rows = []
    for i in range(flat.numel()):
        grads = torch.autograd.grad(flat[i], params, retain_graph=True, allow_unused=True)
        rows.append(
            torch.cat(
                [
                    (g if g is not None else torch.zeros_like(p)).reshape(-1)
                    for g, p in zip(grads, params)
                ]
            )
        )
    return torch.stack(rows), out.detach()


def loss_hessian(loss_fn: nn.Module, out: Tensor, y: Tensor) -> Tensor:
    """Hessian of the (reduced) loss w.r.t. the flattened model output."""
    o = out.detach().clone().requires_grad_(True)
    loss = loss_fn(o, y)
    (g,) = torch.autograd.grad(loss, o, create_graph=True)
    g = g.reshape(-1)
    rows = []
    for i in range(g.numel()):
        (h,) = torch.autograd.grad(g[i], o, retain_graph=True)
        rows.append(h.reshape(-1))
    return torch.stack(rows)


class AutogradGGN:
    """Ground-truth quantities of one problem (model, loss, batch)."""

    def __init__(self, model: nn.Module, loss_fn: nn.Module, x: Tensor, y: Tensor):
        self.model, self.loss_fn, self.x, self.y = model, loss_fn, x, y
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.N = x.shape[0]
        self.J, self.out = output_jacobian(model, x, self.params)  # [N*O, D]
        self.H = loss_hessian(loss_fn, self.out, y)  # [N*O, N*O], block diagonal
        self.O = self.J.shape[0] // self.N
        self._sample_ggn = None

    # -- GGN ---------------------------------------------------------------
    def ggn_batch(self) -> Tensor:
        """``[N, D, D]``: per-sample GGN *including* the reduction factor
        (``autograd.py:102-115``)."""
        if self._sample_ggn is None:
            O = self.O
            mats = []
            for n in range(self.N):
                sl = slice(n * O, (n + 1) * O)
                Jn = self.J[sl]
                mats.append(Jn.t() @ self.H[sl, sl] @ Jn)
            self._sample_ggn = torch.stack(mats)
        return self._sample_ggn

    def ggn(self, subsampling: Optional[Sequence[int]] = None) -> Tensor:
        """Full-batch GGN, or the sum of the selected per-sample GGNs with the
        ``N / len(subsampling)`` rescaling (``autograd.py:95-100,236-238``)."""
        gb = self.ggn_batch()
        if subsampling is None:
            return gb.sum(0)
        return gb[list(subsampling)].sum(0) * (self.N / len(subsampling))

    # -- per-sample gradients ---------------------------------------------
    def batch_grad(self, subsampling: Optional[Sequence[int]] = None) -> Tensor:
        """``[N_grad, D]``; rows are ``d loss / d theta`` restricted to one sample
        (``autograd.py:31-51``), i.e. they carry the ``1/N`` of a mean loss."""
        o = self.out.detach().clone().requires_grad_(True)
        loss = self.loss_fn(o, self.y)
        (g,) = torch.autograd.grad(loss, o)
        g = g.reshape(self.N, -1)
        idx = list(range(self.N)) if subsampling is None else list(subsampling)
        O = self.O
        return torch.stack([g[n] @ self.J[n * O : (n + 1) * O] for n in idx])

    # -- helpers -------------------------------------------------------------
    def group_indices(self, param_groups) -> List[Tensor]:
        """Flat parameter indices of each group (``test/implementation/base.py``)."""
        offsets: Dict[int, Tuple[int, int]] = {}
        start = 0
        for p in self.params:
            offsets[id(p)] = (start, start + p.numel())
            start += p.numel()
        out = []
        for group in param_groups:
            idx = [torch.arange(*offsets[id(p)]) for p in group["params"]]
            out.append(torch.cat(idx) if idx else torch.zeros(0, dtype=torch.long))
        return out

    def directions(self, param_groups, subsampling=None):
        """Eigenpairs of each diagonal GGN block after ``criterion``
        (``autograd.py:221-262``)."""
        ggn = self.ggn(subsampling)
        evals_l, evecs_l = [], []
        for idx, group in zip(self.group_indices(param_groups), param_groups):
            ev, evec = eigh_psd(ggn[idx][:, idx])
            keep = group["criterion"](ev) if "criterion" in group else list(range(ev.numel()))
            evals_l.append(ev[keep])
            evecs_l.append(evec[:, keep])
        return evals_l, evecs_l

    def gammas(self, param_groups, subsampling_ggn=None, subsampling_grad=None):
        """``gamma[n, k] = (N grad_n)^T e_k`` (``autograd.py:123-169``)."""
        _, evecs = self.directions(param_groups, subsampling_ggn)
        g = self.batch_grad(subsampling_grad) * self.N
        return [
            g[:, idx] @ e for idx, e in zip(self.group_indices(param_groups), evecs)
        ], evecs

    def lambdas(self, param_groups, subsampling_ggn=None):
        """``lambda[n, k] = e_k^T (N GGN_n) e_k`` over the GGN samples
        (``autograd.py:171-219``, called with ``lambda_subsampling=subsampling_ggn``
        in ``optim_autograd.py:22-26``)."""
        _, evecs = self.directions(param_groups, subsampling_ggn)
        idx_n = list(range(self.N)) if subsampling_ggn is None else list(subsampling_ggn)
        gb = self.ggn_batch() * self.N
        out = []
        for idx, e in zip(self.group_indices(param_groups), evecs):
            lam = torch.stack(
                [torch.einsum("id,ij,jd->d", e, gb[n][idx][:, idx], e) for n in idx_n]
            )
            out.append(lam)
        return out

    def damped_newton(self, param_groups, subsampling_ggn=None, subsampling_grad=None):
        """``sum_k -gamma_k / (lambda_k + delta_k) e_k`` (``optim_autograd.py:29-59``)."""
        gam, evecs = self.gammas(param_groups, subsampling_ggn, subsampling_grad)
        lam = self.lambdas(param_groups, subsampling_ggn)
        steps = []
        for group, g, l, e in zip(param_groups, gam, lam, evecs):
            delta = group["damping"](None, None, g, l)
            coeff = -g.mean(0) / (l.mean(0) + delta)
            steps.append(e @ coeff)
        return steps

    def ggn_mat_prod(self, param_list, mat: List[Tensor], subsampling=None) -> List[Tensor]:
        """GGN block of ``param_list`` applied to stacked vectors in parameter
        format (``autograd.py:264-312``; note: *no* ``N/len`` rescale there, the
        mean runs over the sub-sampled batch)."""
        (idx,) = self.group_indices([{"params": param_list}])
        gb = self.ggn_batch()
        if subsampling is None:
            ggn = gb.sum(0)
        else:
            ggn = gb[list(subsampling)].sum(0) * (self.N / len(subsampling))
        block = ggn[idx][:, idx]
        flat = torch.cat([m.reshape(m.shape[0], -1) for m in mat], dim=1)
        res = flat @ block.t()
        out, start = [], 0
        for m in mat:
            n = m[0].numel()
            out.append(res[:, start : start + n].reshape(m.shape))
            start += n
        return out
