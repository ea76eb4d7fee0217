"""matrix-free Hessian / GGN operators and Lanczos spectral densities.

Mirrors ``/root/reference/test/hessianfree/test___init__.py:27-110`` (operators against the dense matrices
from autograd for both reductions, several mini-batches; the deterministic-check safeguard) and adds
known-answer checks of the Lanczos quadrature.
"""

import numpy as np
import pytest
import torch
from scipy.sparse.linalg import aslinearoperator
from torch import nn

from oracle.autograd_ggn import AutogradGGN
from vivit_b200.hessianfree import GGNLinearOperator, HessianLinearOperator
from vivit_b200.hessianfree.lanczos import (
    approximate_boundaries,
    approximate_boundaries_abs,
    fast_lanczos,
    lanczos_approximate_log_spectrum,
    lanczos_approximate_spectrum,
)
from vivit_b200.hessianfree.utils import LowRank, Projector


def _problem(reduction):
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(6, 5), nn.Sigmoid(), nn.Linear(5, 4), nn.Tanh(), nn.Linear(4, 3)).double()
    loss = nn.CrossEntropyLoss(reduction=reduction)
    batches = [(torch.rand(n, 6, dtype=torch.float64), torch.randint(0, 3, (n,))) for n in (4, 2, 3)]
    return model, loss, batches


def _dense_hessian(model, loss, X, y):
    params = list(model.parameters())
    out = loss(model(X), y)
    g = torch.autograd.grad(out, params, create_graph=True)
    flat = torch.cat([t.reshape(-1) for t in g])
    rows = [torch.cat([h.reshape(-1) for h in torch.autograd.grad(flat[i], params, retain_graph=True)])
            for i in range(flat.numel())]
    return torch.stack(rows)


@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_operators_match_dense_matrices(reduction):
    model, loss, batches = _problem(reduction)
    X, y = torch.cat([b[0] for b in batches]), torch.cat([b[1] for b in batches])
    ggn = AutogradGGN(model, loss, X, y).ggn().numpy()
    hess = _dense_hessian(model, loss, X, y).numpy()
    G = GGNLinearOperator(model, loss, batches, torch.device("cpu"), dtype=np.float64)
    H = HessianLinearOperator(model, loss, batches, torch.device("cpu"), dtype=np.float64)
    D = G.shape[0]
    assert G.shape == H.shape == (D, D) == ggn.shape
    rng = np.random.default_rng(0)
    V = rng.standard_normal((D, 3))
    assert np.allclose(G @ V, ggn @ V, rtol=1e-9, atol=1e-12)
    assert np.allclose(H @ V, hess @ V, rtol=1e-9, atol=1e-12)
    v = torch.from_numpy(V[:, 0])
    assert np.allclose(G.matvec_torch(v).numpy(), ggn @ V[:, 0], rtol=1e-9, atol=1e-12)
    grad, value = G.gradient_and_loss()
    full = loss(model(X), y)
    want = torch.autograd.grad(full, list(model.parameters()))
    assert torch.allclose(value, full.detach().reshape(1).float(), rtol=1e-6)
    for g, w in zip(grad, want):
        assert torch.allclose(g, w, rtol=1e-9, atol=1e-12)


def test_ggn_spectrum_matches_the_gram_path():
    """The matrix-free operator and the low-rank Gram path see the same non-zero GGN spectrum."""
    from oracle import reference_path as ref

    model, loss, batches = _problem("mean")
    X, y = torch.cat([b[0] for b in batches]), torch.cat([b[1] for b in batches])
    G = GGNLinearOperator(model, loss, [(X, y)], torch.device("cpu"), dtype=np.float64)
    dense = G @ np.eye(G.shape[0])
    top = np.linalg.eigvalsh((dense + dense.T) / 2)[-5:]
    (gram_evals,) = ref.eigvalsh(model, loss, X, y, [{"params": list(model.parameters())}])
    assert np.allclose(top, gram_evals[-5:].numpy(), rtol=1e-8, atol=1e-12)


def test_deterministic_check_catches_dropout():
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(6, 8), nn.Dropout(0.5), nn.Linear(8, 3)).double().train()
    batches = [(torch.rand(5, 6, dtype=torch.float64), torch.randint(0, 3, (5,)))]
    with pytest.raises(RuntimeError):
        GGNLinearOperator(model, nn.CrossEntropyLoss(), batches, torch.device("cpu"), dtype=np.float64)
    GGNLinearOperator(model.eval(), nn.CrossEntropyLoss(), batches, torch.device("cpu"), dtype=np.float64)
    with pytest.raises(ValueError):
        GGNLinearOperator(model, nn.CrossEntropyLoss(reduction="none"), batches, torch.device("cpu"),
                          check_deterministic=False) @ np.ones(8 * 6 + 8 + 3 * 8 + 3, dtype=np.float32)


def test_lanczos_quadrature_known_answers():
    rng = np.random.default_rng(1)
    evals = np.r_[np.linspace(0.5, 1.0, 30), 5.0, 9.0]
    Q, _ = np.linalg.qr(rng.standard_normal((32, 32)))
    A = aslinearoperator(Q @ np.diag(evals) @ Q.T)
    np.random.seed(0)
    ritz, vecs = fast_lanczos(A, 32)
    # no re-orthogonalisation: converged Ritz values may appear twice ("ghosts"), but both outliers are found
    assert abs(ritz[-1] - 9.0) < 1e-6 and np.abs(ritz - 5.0).min() < 1e-6
    assert np.allclose((vecs[0] ** 2).sum(), 1.0)
    ritz2, _ = fast_lanczos(A, 20, use_eigh_tridiagonal=True)
    assert abs(ritz2[-1] - 9.0) < 1e-6
    lo, hi = approximate_boundaries(A, tol=1e-6)
    assert abs(lo - 0.5) < 1e-4 and abs(hi - 9.0) < 1e-4
    lo_abs, hi_abs = approximate_boundaries_abs(A, tol=1e-6)
    assert abs(lo_abs - 0.5) < 1e-4 and abs(hi_abs - 9.0) < 1e-4
    grid, density = lanczos_approximate_spectrum(A, 32, num_points=2048, num_repeats=4, boundaries=(0.5, 9.0))
    assert grid[0] < 0.5 and grid[-1] > 9.0 and density.min() >= 0
    assert abs(np.trapezoid(density, grid) - 1.0) < 2e-2
    bulk = np.trapezoid(density[grid < 3.0], grid[grid < 3.0])
    assert abs(bulk - 30 / 32) < 5e-2  # 30 of the 32 eigenvalues sit in [0.5, 1]
    lgrid, ldensity = lanczos_approximate_log_spectrum(A, 32, num_points=2048, num_repeats=4, boundaries=(0.5, 9.0))
    assert ldensity.min() >= 0 and abs(np.trapezoid(ldensity, lgrid) - 1.0) < 5e-2


def test_low_rank_operators():
    rng = np.random.default_rng(2)
    A, c, x = rng.standard_normal((7, 3)), rng.standard_normal(3), rng.standard_normal(7)
    assert np.allclose(LowRank(c, A) @ x, (A * c) @ A.T @ x)
    Q, _ = np.linalg.qr(A)
    P = Projector(Q)
    assert np.allclose(P @ (P @ x), P @ x) and np.allclose(P @ Q[:, 0], Q[:, 0])


def test_device_recurrence_on_operator():
    """``fast_lanczos`` on a curvature operator takes the ``matvec_torch`` path (on CPU here)."""
    model, loss, batches = _problem("mean")
    G = GGNLinearOperator(model, loss, batches, torch.device("cpu"), dtype=np.float64)
    dense = G @ np.eye(G.shape[0])
    np.random.seed(0)
    ritz, vecs = fast_lanczos(G, 25)
    assert abs(ritz[-1] - np.linalg.eigvalsh((dense + dense.T) / 2)[-1]) < 1e-8 * max(1.0, abs(ritz[-1]))
    assert np.allclose((vecs[0] ** 2).sum(), 1.0)


def test_tridiagonal_decomposition_through_the_library_solver(monkeypatch):
    """The GPU branch of ``fast_lanczos`` (dense Lanczos matrix -> ``kernels.syevj``), exercised on CPU through
    the test double: an indefinite matrix is shifted to positive semi-definite and shifted back."""
    import tests._torch_kernels as double
    from scipy.linalg import eigh_tridiagonal
    from vivit_b200.hessianfree.lanczos import _eigh_tridiagonal_device

    double.install(monkeypatch)
    rng = np.random.default_rng(3)
    alphas, betas = rng.standard_normal(12) - 0.5, rng.standard_normal(11)
    evals, evecs = _eigh_tridiagonal_device(torch.from_numpy(alphas), torch.from_numpy(betas))
    want, wvecs = eigh_tridiagonal(alphas, betas)
    assert want[0] < 0 < want[-1]
    assert np.allclose(evals, want, rtol=1e-10, atol=1e-12)
    assert np.allclose(np.abs(evecs[0]), np.abs(wvecs[0]), rtol=1e-8, atol=1e-10)
