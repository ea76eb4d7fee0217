"""the Gram extension hooks of ``vivit/extensions/hooks.py`` -- oracle pinned against
autograd, host logic (hook protocol, keyword arguments) through the test double on CPU.

Mirrors ``/root/reference/test/extensions/firstorder/batch_grad/test_gram_batch_grad.py:21-171`` and
``test/extensions/secondorder/sqrt_ggn/test_gram_sqrt_ggn.py:17-109``.
"""

import pytest
import torch

import tests._torch_kernels as double
from oracle import reference_path as ref
from oracle.autograd_ggn import AutogradGGN
from tests.problems import IDS, PROBLEM_SUM, PROBLEMS
from tests.test_host_cpu import close, run_backward

ALL = PROBLEMS + [PROBLEM_SUM]
ALL_IDS = IDS + ["mlp-ce-sum"]


@pytest.fixture(autouse=True)
def torch_kernels(monkeypatch):
    double.install(monkeypatch)


def nonzero(evals, rtol=1e-5, atol=1e-7):
    return evals[~torch.isclose(evals, torch.zeros_like(evals), rtol=rtol, atol=atol)]


# ---- the oracle against the autograd ground truth ---------------------------------------------


@pytest.mark.parametrize("problem", ALL, ids=ALL_IDS)
def test_oracle_gram_batch_grad_vs_autograd(problem):
    """``test_gram_batch_grad.py:21-75``: Gram of per-sample gradients (``autograd.py:21-28``)
    and the spectrum of the (centred) gradient covariance."""
    model, loss, x, y = problem.make(torch.float64)
    g = AutogradGGN(model, loss, x, y).batch_grad()  # [N, D]
    got, _ = ref.gram_batch_grad(model, loss, x, y)
    close(got, g @ g.t())
    gc = g - g.mean(0)
    got_c, _ = ref.gram_batch_grad(model, loss, x, y, center=True)
    close(got_c, gc @ gc.t())
    for gram, rows in ((got, g), (got_c, gc)):
        cov = nonzero(torch.linalg.eigvalsh(rows.t() @ rows))
        ev = nonzero(torch.linalg.eigvalsh(gram))
        n = min(cov.numel(), ev.numel())
        close(ev[-n:], cov[-n:], 1e-5, 1e-7)
    flat = torch.cat([c.flatten(1) for c in ref.centered_batch_grad(model, loss, x, y)], 1)
    close(flat, gc, 1e-5, 1e-7)


@pytest.mark.parametrize("problem", ALL, ids=ALL_IDS)
def test_oracle_gram_sqrt_ggn_spectrum_vs_autograd(problem):
    """``test_gram_sqrt_ggn.py:17-31``: non-zero spectrum of the GGN Gram matrix == that of the GGN."""
    model, loss, x, y = problem.make(torch.float64)
    gram, _ = ref.gram_sqrt_ggn(model, loss, x, y)
    ggn_ev = nonzero(torch.linalg.eigvalsh(AutogradGGN(model, loss, x, y).ggn()), atol=2e-6)
    gram_ev = nonzero(torch.linalg.eigvalsh(gram), atol=2e-6)
    close(gram_ev, ggn_ev, 1e-5, 1e-7)


# ---- the hooks (host logic; kernels replaced by the test double) ---------------------------------


@pytest.mark.parametrize("lazy", [False, True], ids=["tensor", "factor"])
@pytest.mark.parametrize("problem", ALL, ids=ALL_IDS)
def test_gram_batch_grad_hooks_match_oracle(problem, lazy):
    from vivit_b200 import BatchGrad
    from vivit_b200.extensions.hooks import CenteredBatchGrad, CenteredGramBatchGrad, GramBatchGrad

    model, loss, x, y = problem.make(torch.float64)
    for center, cls in ((False, GramBatchGrad), (True, CenteredGramBatchGrad)):
        hook = cls()
        run_backward(model, loss, x, y, [BatchGrad(lazy=lazy)], hook)
        want, _ = ref.gram_batch_grad(model, loss, x, y, center=center)
        close(hook.get_result(), want)
        assert hook.get_result().shape == (x.shape[0], x.shape[0])
    hook = CenteredBatchGrad()
    run_backward(model, loss, x, y, [BatchGrad(lazy=lazy)], hook)
    for p, want in zip(model.parameters(), ref.centered_batch_grad(model, loss, x, y)):
        close(p.centered_grad_batch, want)
        assert p.centered_grad_batch.shape == (x.shape[0], *p.shape)


@pytest.mark.parametrize("lazy", [False, True], ids=["tensor", "factor"])
@pytest.mark.parametrize("problem", ALL, ids=ALL_IDS)
def test_gram_sqrt_ggn_hooks_match_oracle(problem, lazy):
    from vivit_b200 import SqrtGGNExact, SqrtGGNMC
    from vivit_b200.extensions.hooks import GramSqrtGGNExact, GramSqrtGGNMC

    model, loss, x, y = problem.make(torch.float64)
    hook = GramSqrtGGNExact()
    run_backward(model, loss, x, y, [SqrtGGNExact(lazy=lazy)], hook)
    want, _ = ref.gram_sqrt_ggn(model, loss, x, y)
    close(hook.get_result(), want)
    if isinstance(loss, torch.nn.CrossEntropyLoss):
        M, N = 3, x.shape[0]
        torch.manual_seed(1)
        with torch.no_grad():
            ids = ref.sample_ce_classes(model(x), None, M)
        ext = SqrtGGNMC(mc_samples=M, lazy=lazy)
        ext.mc_state = ids
        hook = GramSqrtGGNMC()
        run_backward(model, loss, x, y, [ext], hook)
        want, _ = ref.gram_sqrt_ggn(model, loss, x, y, mc_samples=M, mc_state=ids)
        assert hook.get_result().shape == (M * N, M * N)
        close(hook.get_result(), want)


@pytest.mark.parametrize("free", [True, False], ids=["free", "keep"])
@pytest.mark.parametrize("layerwise", [True, False], ids=["layerwise", "total-only"])
@pytest.mark.parametrize("problem", PROBLEMS[:2], ids=IDS[:2])
def test_keyword_arguments(problem, layerwise, free):
    """``test_gram_batch_grad.py:104-171``, ``test_gram_sqrt_ggn.py:67-109``: ``free_*`` deletes the
    source savefield, ``layerwise`` keeps the per-parameter matrices (else the savefield is None)."""
    from vivit_b200 import BatchGrad, SqrtGGNExact
    from vivit_b200.extensions.hooks import CenteredGramBatchGrad, GramBatchGrad, GramSqrtGGNExact

    model, loss, x, y = problem.make(torch.float64)
    cases = [
        (GramBatchGrad, "free_grad_batch", BatchGrad, "grad_batch", "gram_grad_batch",
         lambda: ref.gram_batch_grad(model, loss, x, y)),
        (CenteredGramBatchGrad, "free_grad_batch", BatchGrad, "grad_batch", "centered_gram_grad_batch",
         lambda: ref.gram_batch_grad(model, loss, x, y, center=True)),
        (GramSqrtGGNExact, "free_sqrt_ggn", SqrtGGNExact, "sqrt_ggn_exact", "gram_sqrt_ggn_exact",
         lambda: ref.gram_sqrt_ggn(model, loss, x, y)),
    ]
    for cls, free_kw, ext, source, savefield, oracle in cases:
        hook = cls(layerwise=layerwise, **{free_kw: free})
        run_backward(model, loss, x, y, [ext()], hook)
        total, layers = oracle()
        close(hook.get_result(), total)
        for p in model.parameters():
            assert hasattr(p, source) != free
            if layerwise:
                close(getattr(p, savefield), layers[id(p)])
            else:
                assert getattr(p, savefield) is None
            for name in (source, savefield):
                if hasattr(p, name):
                    delattr(p, name)


def test_centered_gram_centres_the_stored_gradients_in_place():
    """``gram_batch_grad.py:88-89``: ``grad_batch -= grad_batch.mean(0)`` is visible to the caller."""
    from vivit_b200 import BatchGrad
    from vivit_b200.extensions.hooks import CenteredGramBatchGrad

    model, loss, x, y = PROBLEMS[0].make(torch.float64)
    run_backward(model, loss, x, y, [BatchGrad()], CenteredGramBatchGrad())
    for p, want in zip(model.parameters(), ref.centered_batch_grad(model, loss, x, y)):
        close(p.grad_batch, want)


def test_hook_without_savefield_must_not_return_values():
    from vivit_b200.utils.hooks import ParameterHook

    class Bad(ParameterHook):
        def param_hook(self, param):
            return 1

    model, loss, x, y = PROBLEMS[0].make(torch.float64)
    with pytest.raises(ValueError):
        run_backward(model, loss, x, y, [], Bad())
