"""Pin the oracle: ``oracle/reference_path.py`` against the autograd ground truth.

Mirrors the reference's own hot-path tests and tolerances
(``/root/reference/test/linalg/test_eigvalsh.py:24-63``, ``test_eigh.py:25-155``,
``test/optim/test_directional_derivatives.py:27-72``,
``test/optim/test_directional_damped_newton.py:27-74``,
``test/extensions/secondorder/vivit/test_vivit_ggn.py:20-113``).
"""

import pytest
import torch

from oracle import reference_path as ref
from oracle.autograd_ggn import AutogradGGN
from tests.problems import (
    GROUPING_IDS,
    GROUPINGS,
    IDS,
    PROBLEM_SUM,
    PROBLEMS,
    constant_damping,
    keep_all,
    keep_nonzero,
    make_top_k,
)

SUBS = [None, [1, 0]]
SUB_IDS = ["full", "sub10"]


def _close(a, b, rtol, atol):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a - b).abs().max()


@pytest.mark.parametrize("grouping", GROUPINGS, ids=GROUPING_IDS)
@pytest.mark.parametrize("sub", SUBS, ids=SUB_IDS)
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_eigvalsh(problem, sub, grouping):
    model, loss, x, y = problem.make(torch.float64)
    groups = grouping(model, criterion=keep_all)
    got = ref.eigvalsh(model, loss, x, y, groups, subsampling=sub)
    want, _ = AutogradGGN(model, loss, x, y).directions(groups, sub)
    for g, w in zip(got, want):
        n = min(g.numel(), w.numel())
        _close(g[-n:], w[-n:], 1e-8, 1e-12)


@pytest.mark.parametrize("dtype,rtol,atol", [(torch.float64, 1e-7, 1e-10), (torch.float32, 5e-4, 1e-5)])
@pytest.mark.parametrize("grouping", GROUPINGS, ids=GROUPING_IDS)
@pytest.mark.parametrize("sub", SUBS, ids=SUB_IDS)
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_eigh(problem, sub, grouping, dtype, rtol, atol):
    model, loss, x, y = problem.make(dtype)
    groups = grouping(model, criterion=keep_nonzero)
    got = ref.eigh(model, loss, x, y, groups, subsampling=sub)
    truth = AutogradGGN(model, loss, x, y)
    want_evals, _ = truth.directions(groups, sub)
    for (evals, evecs), w, group in zip(got, want_evals, groups):
        n = min(evals.numel(), w.numel())
        _close(evals[-n:], w[-n:], 1e-4, 5e-6)
        # GGN e = lambda e   (test_eigh.py:118-133)
        Ge = truth.ggn_mat_prod(group["params"], evecs, sub)
        for a, e in zip(Ge, evecs):
            _close(a, torch.einsum("i,i...->i...", evals, e), rtol, atol)
        # orthonormal (test_eigh.py:135-143)
        gram = sum(torch.einsum("i...,j...->ij", e, e) for e in evecs)
        _close(gram, torch.eye(evals.numel(), dtype=dtype), 1e-3, 2e-4)


@pytest.mark.parametrize("k", [1, 10])
@pytest.mark.parametrize("sub_ggn", [None, [0, 1]], ids=["ggn-full", "ggn-01"])
@pytest.mark.parametrize("sub_grad", [None, [0, 1]], ids=["grad-full", "grad-01"])
@pytest.mark.parametrize("grouping", GROUPINGS, ids=GROUPING_IDS)
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_directional_derivatives_and_newton(problem, grouping, sub_grad, sub_ggn, k):
    model, loss, x, y = problem.make(torch.float64)
    groups = grouping(model, criterion=make_top_k(k), damping=constant_damping(1.0))
    truth = AutogradGGN(model, loss, x, y)
    want_g, _ = truth.gammas(groups, sub_ggn, sub_grad)
    want_l = truth.lambdas(groups, sub_ggn)
    got = ref.directional_derivatives(
        model, loss, x, y, groups, subsampling_grad=sub_grad, subsampling_ggn=sub_ggn
    )
    for (g, l), wg, wl in zip(got, want_g, want_l):
        _close(g.abs(), wg.abs(), 1e-5, 1e-8)  # sign of an eigenvector is free
        _close(l, wl, 1e-5, 1e-8)
    steps = ref.directional_damped_newton(
        model, loss, x, y, groups, subsampling_grad=sub_grad, subsampling_ggn=sub_ggn
    )
    want_steps = truth.damped_newton(groups, sub_ggn, sub_grad)
    for s, w in zip(steps, want_steps):
        _close(torch.cat([t.flatten() for t in s]), w, 1e-5, 1e-8)


@pytest.mark.parametrize("sub", [None, [0, 0, 1, 0, 1]], ids=["full", "repeat"])
@pytest.mark.parametrize("problem", PROBLEMS + [PROBLEM_SUM], ids=IDS + ["mlp-ce-sum"])
def test_vivit_closures_mat_prod(problem, sub):
    """``V (V^T M) == G M`` and structured == dense Gram (test_vivit_ggn.py:20-52)."""
    model, loss, x, y = problem.make(torch.float64)
    if sub is not None and max(sub) >= x.shape[0]:
        pytest.skip("batch too small")
    sweep = ref.backward_sweep(model, loss, x, y, subsampling_ggn=sub, want_vivit=True, want_sqrt_ggn=True)
    params = list(model.parameters())
    torch.manual_seed(1)
    mat = [torch.rand(3, *p.shape, dtype=torch.float64) for p in params]
    vt = sum(sweep.vivit[id(p)]["V_t_mat_prod"](m) for p, m in zip(params, mat))
    got = [sweep.vivit[id(p)]["V_mat_prod"](vt) for p in params]
    # ground truth: GGN of the sub-sampled mini-batch *without* N/len rescale and,
    # for repeated indices, with multiplicity
    truth = AutogradGGN(model, loss, x, y)
    gb = truth.ggn_batch()
    ggn = gb.sum(0) if sub is None else gb[sub].sum(0)
    flat = torch.cat([m.reshape(3, -1) for m in mat], 1) @ ggn.t()
    got_flat = torch.cat([g.reshape(3, -1) for g in got], 1)
    _close(got_flat, flat, 1e-7, 1e-10)
    for p in params:
        dense = ref.pairwise_dot(sweep.sqrt_ggn[id(p)], 2)
        _close(sweep.vivit[id(p)]["gram_mat"](), dense, 1e-9, 1e-12)


def test_mc_factor_converges():
    """E[(p - e_y)(p - e_y)^T] = diag(p) - p p^T (test_vivit_ggn.py:79-113, loosened)."""
    torch.manual_seed(0)
    logits = torch.randn(3, 5, dtype=torch.float64)
    exact = ref.sqrt_hessian_ce(logits, None)
    h_exact = torch.einsum("vnc,vnd->ncd", exact, exact)
    ids = ref.sample_ce_classes(logits, None, 100000)
    mc = ref.sqrt_hessian_ce_sampled(logits, None, ids)
    h_mc = torch.einsum("vnc,vnd->ncd", mc, mc)
    _close(h_mc, h_exact, 1e-1, 1e-3)


def test_remove_zero_evals():
    ev = torch.tensor([0.0, 1e-8, 1e-3, 2.0])
    evals, evecs = ref.remove_zero_evals(ev, torch.eye(4))
    assert evals.tolist() == [ev[2].item(), 2.0] and evecs.shape == (4, 2)
