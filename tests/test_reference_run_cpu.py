"""Pinning against outputs of the reference's OWN code.

``tests/golden/reference_run.pt`` holds what f-dangel/vivit's ``EigvalshComputation``,
``EighComputation``, ``DirectionalDerivativesComputation`` and ``DirectionalDampedNewtonComputation``
returned when their unmodified hooks ran in the build container on every fixture of
``tests/problems.py`` (``tests/golden/make_reference_run.py``: BackPACK's per-parameter tensors by
autograd, import stubs for the BackPACK names, ``Tensor.symeig`` shimmed to ``linalg.eigh``).

* the oracle restatement (``oracle/reference_path.py``) must reproduce them;
* so must the shipped host code (kernel layer replaced by the test double on this CPU-only machine);
* and the autograd ground truth of ``tests/golden/ground_truth.pt`` must agree with them;
* the same for the extension hooks of ``vivit.extensions.hooks`` (``"__gram_hooks__"``) and for
  ``vivit/hessianfree/lanczos.py`` / ``utils.py`` (``"__lanczos__"``: fixed matrix, fixed numpy seed,
  explicit boundaries -- the boundary estimate itself is an ``eigsh`` run with tolerance 1e-2).

The reference factorises the loss Hessian its own way; the generator uses a symmetric ``eigh`` factor.
Eigenvalues, directional derivatives and Newton steps do not depend on that choice, eigenvectors are
compared through the projector onto the kept eigenspace, and the number of (numerically zero)
trailing Gram eigenvalues may differ, so spectra are compared from the top.
"""

import os

import pytest
import torch

import tests._torch_kernels as double
from oracle import reference_path as ref
from tests.problems import (
    GROUPING_IDS,
    GROUPINGS,
    IDS,
    PROBLEMS,
    constant_damping,
    keep_nonzero,
    make_top_k,
)
from tests.test_host_cpu import run_backward

RUN = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_run.pt"))
TRUTH = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ground_truth.pt"))
SUBS = [("full", None), ("sub10", [1, 0])]


def close(got, want, tol=1e-9, what=""):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if want.numel():
        scale = max(want.abs().max().item(), 1e-300)
        err = (got - want).abs().max().item() / scale
        assert err <= tol, f"{what}: {err:.3e} relative to the largest entry"


def top(got, want, what):
    n = min(got.numel(), want.numel())
    close(got[-n:], want[-n:], what=what)
    assert got[: got.numel() - n].abs().max().item() <= 1e-12 if got.numel() > n else True
    assert want[: want.numel() - n].abs().max().item() <= 1e-12 if want.numel() > n else True


def projector(evecs_flat):
    return evecs_flat.t() @ evecs_flat


def test_every_fixture_was_run_by_the_reference():
    cases = {k for k in RUN if isinstance(k, tuple)}
    assert cases == {(p.name, g, s) for p in PROBLEMS for g in GROUPING_IDS for s, _ in SUBS}
    assert RUN["__meta__"]["batch_sizes"] == {p.name: p.make()[2].shape[0] for p in PROBLEMS}


@pytest.mark.parametrize("grouping,gname", list(zip(GROUPINGS, GROUPING_IDS)), ids=GROUPING_IDS)
@pytest.mark.parametrize("sname,sub", SUBS, ids=[s for s, _ in SUBS])
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_oracle_reproduces_the_reference_run(problem, sname, sub, grouping, gname):
    want = RUN[(problem.name, gname, sname)]
    model, loss, x, y = problem.make(torch.float64)

    for got, w in zip(ref.eigvalsh(model, loss, x, y, grouping(model), subsampling=sub), want["eigvalsh"]):
        top(got, w, "eigvalsh")

    groups = grouping(model, criterion=keep_nonzero)
    for (evals, evecs), w_evals, w_evecs in zip(
        ref.eigh(model, loss, x, y, groups, subsampling=sub), want["eigh_evals"], want["eigh_evecs"]
    ):
        close(evals, w_evals, what="eigh evals")
        flat = torch.cat([e.flatten(1) for e in evecs], 1)
        assert (projector(flat) - projector(w_evecs)).abs().max().item() <= 1e-9

    groups = grouping(model, criterion=make_top_k(10), damping=constant_damping(1.0))
    derivs = ref.directional_derivatives(model, loss, x, y, groups, sub, sub)
    for (gam, lam), wg, wl in zip(derivs, want["gammas_abs"], want["lambdas"]):
        close(gam.abs(), wg, what="gammas")
        close(lam, wl, what="lambdas")
    steps = ref.directional_damped_newton(model, loss, x, y, groups, sub, sub)
    for step, w in zip(steps, want["newton"]):
        close(torch.cat([s.flatten() for s in step]), w, what="newton")


@pytest.mark.parametrize("grouping,gname", list(zip(GROUPINGS, GROUPING_IDS)), ids=GROUPING_IDS)
@pytest.mark.parametrize("sname,sub", SUBS, ids=[s for s, _ in SUBS])
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_host_code_reproduces_the_reference_run(problem, sname, sub, grouping, gname, monkeypatch):
    from vivit_b200 import (
        DirectionalDampedNewtonComputation,
        DirectionalDerivativesComputation,
        EighComputation,
        EigvalshComputation,
    )

    double.install(monkeypatch)
    want = RUN[(problem.name, gname, sname)]
    model, loss, x, y = problem.make(torch.float64)

    groups = grouping(model)
    comp = EigvalshComputation(subsampling=sub)
    run_backward(model, loss, x, y, [comp.get_extension()], comp.get_extension_hook(groups))
    for g, w in zip(groups, want["eigvalsh"]):
        top(comp.get_result(g), w, "eigvalsh")

    groups = grouping(model, criterion=keep_nonzero)
    comp = EighComputation(subsampling=sub, warn_small_eigvals=0.0)
    run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))
    for g, w_evals, w_evecs in zip(groups, want["eigh_evals"], want["eigh_evecs"]):
        evals, evecs = comp.get_result(g)
        close(evals, w_evals, what="eigh evals")
        flat = torch.cat([e.flatten(1) for e in evecs], 1)
        assert (projector(flat) - projector(w_evecs)).abs().max().item() <= 1e-9

    groups = grouping(model, criterion=make_top_k(10), damping=constant_damping(1.0))
    comp = DirectionalDerivativesComputation(subsampling_grad=sub, subsampling_ggn=sub, warn_small_eigvals=0.0)
    run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))
    for g, wg, wl in zip(groups, want["gammas_abs"], want["lambdas"]):
        gam, lam = comp.get_result(g)
        close(gam.abs(), wg, what="gammas")
        close(lam, wl, what="lambdas")
    comp = DirectionalDampedNewtonComputation(subsampling_grad=sub, subsampling_ggn=sub, warn_small_eigvals=0.0)
    run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))
    for g, w in zip(groups, want["newton"]):
        close(torch.cat([s.flatten() for s in comp.get_result(g)]), w, what="newton")


def test_autograd_ground_truth_agrees_with_the_reference_run():
    for key, want in RUN.items():
        if not isinstance(key, tuple):
            continue
        truth = TRUTH[key]
        for name in ("gammas_abs", "lambdas", "newton"):
            for got, w in zip(truth[name], want[name]):
                close(got, w, what=f"{key} {name}")
        for got, w in zip(truth["evals_all"], want["eigvalsh"]):
            n = min(got.numel(), w.numel())  # GGN [D, D] against Gram [R, R]: the non-trivial part is shared
            close(got[-n:], w[-n:], tol=1e-8, what=f"{key} spectrum")


# ---- extension hooks ------------------------------------------------------------------


def check_layerwise(got, want, want_total, what):
    """Layer-wise Gram matrices.  Upstream quirk, kept out of this code on purpose: the reference's
    ``_update_result`` (``gram_batch_grad.py:113-118``) adopts the FIRST visited parameter's matrix as its
    accumulator and then adds into it in place, so that parameter's layer-wise savefield ends up holding
    the total over all parameters (its tests only check that the savefield exists).  Here every parameter
    keeps its own matrix: equal to the reference's wherever the reference's is not that alias, and summing
    to the reference's total."""
    aliased = [i for i, w in enumerate(want) if torch.equal(w, want_total)]
    assert len(aliased) >= 1
    for i, (g, w) in enumerate(zip(got, want)):
        if i not in aliased or len(want) == 1:
            close(g, w, what=what + " layerwise")
    close(sum(got), want_total, what=what + " layerwise sum")


@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_oracle_gram_hooks_reproduce_the_reference_run(problem):
    want = RUN["__gram_hooks__"][problem.name]
    model, loss, x, y = problem.make(torch.float64)
    for name, center in (("gram", False), ("centered_gram", True)):
        total, layers = ref.gram_batch_grad(model, loss, x, y, center=center)
        close(total, want[name], what=name)
        check_layerwise([layers[id(p)] for p in model.parameters()], want[name + "_layerwise"], want[name], name)
    for got, w in zip(ref.centered_batch_grad(model, loss, x, y), want["centered_grad_batch"]):
        close(got, w, what="centered_grad_batch")
    gram, _ = ref.gram_sqrt_ggn(model, loss, x, y)
    top(torch.linalg.eigvalsh(gram), want["gram_sqrt_ggn_evals"], "gram_sqrt_ggn spectrum")


@pytest.mark.parametrize("lazy", [False, True], ids=["tensor", "factor"])
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_host_gram_hooks_reproduce_the_reference_run(problem, lazy, monkeypatch):
    from vivit_b200 import BatchGrad, SqrtGGNExact
    from vivit_b200.extensions.hooks import CenteredBatchGrad, CenteredGramBatchGrad, GramBatchGrad, GramSqrtGGNExact

    double.install(monkeypatch)
    want = RUN["__gram_hooks__"][problem.name]
    model, loss, x, y = problem.make(torch.float64)
    for name, cls in (("gram", GramBatchGrad), ("centered_gram", CenteredGramBatchGrad)):
        hook = cls(layerwise=True, free_grad_batch=True)
        run_backward(model, loss, x, y, [BatchGrad(lazy=lazy)], hook)
        close(hook.get_result(), want[name], what=name)
        check_layerwise(
            [getattr(p, hook.savefield) for p in model.parameters()], want[name + "_layerwise"], want[name], name
        )
        for p in model.parameters():
            delattr(p, hook.savefield)
    hook = CenteredBatchGrad()
    run_backward(model, loss, x, y, [BatchGrad(lazy=lazy)], hook)
    for p, w in zip(model.parameters(), want["centered_grad_batch"]):
        close(p.centered_grad_batch, w, what="centered_grad_batch")
    hook = GramSqrtGGNExact()
    run_backward(model, loss, x, y, [SqrtGGNExact(lazy=lazy)], hook)
    top(torch.linalg.eigvalsh(hook.get_result()), want["gram_sqrt_ggn_evals"], "gram_sqrt_ggn spectrum")


# ---- hessianfree: Lanczos quadrature and low-rank operators ----------------------------


def test_lanczos_functions_reproduce_the_reference_run():
    import numpy as np
    from scipy.sparse.linalg import aslinearoperator

    from vivit_b200.hessianfree import lanczos
    from vivit_b200.hessianfree.utils import LowRank, Projector

    want = RUN["__lanczos__"]
    op = aslinearoperator(want["A"].numpy())
    for name in ("fast_lanczos", "fast_lanczos_tridiagonal", "lanczos_approximate_spectrum",
                 "lanczos_approximate_log_spectrum"):
        fn = getattr(lanczos, "fast_lanczos" if name.startswith("fast_lanczos") else name)
        np.random.seed(want["seed"])
        got = fn(op, **want[name]["kwargs"])
        for g, w in zip(got, want[name]["result"]):
            g = torch.from_numpy(np.asarray(g, dtype=np.float64))
            if name.startswith("fast_lanczos") and g.dim() == 2:  # eigenvectors: up to sign
                g, w = g.abs(), w.abs()
            close(g, w, tol=1e-9, what=name)
    lr = want["low_rank"]
    c, A, x = lr["c"].numpy(), lr["A"].numpy(), lr["x"].numpy()
    close(torch.from_numpy(LowRank(c, A) @ x), lr["LowRank"], what="LowRank")
    close(torch.from_numpy(Projector(A) @ x), lr["Projector"], what="Projector")


# ---- structured closures of a Linear weight ------------------------------------------


def test_linear_weight_closures_reproduce_the_reference_run(monkeypatch):
    """``ViViTGGNLinear.weight`` (``linear.py:29-81``) run by the reference on seeded tensors: the oracle's
    closures and the product's ``LinearWeightFactor`` (kernels replaced by the test double here; the GPU
    twin is ``test_parity_gpu.py::test_linear_weight_factor_against_the_reference_run``)."""
    from vivit_b200.factors import LinearWeightFactor

    double.install(monkeypatch)
    for case in RUN["__linear_closures__"]:
        s, x, sub = case["s"], case["input0"], case["subsampling"]
        z = x if sub is None else x[sub]
        fns = ref._linear_weight_closures(s, z)
        factor = LinearWeightFactor(s.contiguous(), z.contiguous())
        mat_v, mat_vt = case["mat_v"], case["mat_vt"]
        for got in (fns["gram_mat"](), factor.gram_mat()):
            close(got.reshape(case["gram_mat"].shape), case["gram_mat"], what="gram_mat")
        for got in (fns["V_mat_prod"](mat_v), factor.backtransform(mat_v.reshape(mat_v.shape[0], -1), None)):
            close(got.reshape(case["V_mat_prod"].shape), case["V_mat_prod"], what="V_mat_prod")
        for got in (fns["V_t_mat_prod"](mat_vt), factor.vt_mat_prod(mat_vt)):
            close(got.reshape(case["V_t_mat_prod"].shape), case["V_t_mat_prod"], what="V_t_mat_prod")


# ---- vivit/utils/eig.py: shifted / filtered decompositions ----------------------------


def _same_pairs(got, want, what):
    (evals, evecs), (w_evals, w_evecs) = got, want
    close(evals, w_evals, what=what + " evals")
    assert evecs.shape == w_evecs.shape, (what, evecs.shape, w_evecs.shape)
    if w_evecs.numel():
        # columns up to sign where the eigenvalue is simple; the numerically-zero ones span a degenerate
        # space, compared through its projector
        simple = w_evals.abs() > 1e-8 * w_evals.abs().max()
        a, b = evecs[:, simple], w_evecs[:, simple]
        close(a * torch.sign((a * b).sum(0)), b, tol=1e-8, what=what + " evecs")
        a, b = evecs[:, ~simple], w_evecs[:, ~simple]
        assert (a @ a.t() - b @ b.t()).abs().max().item() <= 1e-8, what + " null space"


def test_eig_utils_reproduce_the_reference_run(monkeypatch):
    from vivit_b200.utils import eig

    double.install(monkeypatch)
    run = RUN["__eig_utils__"]
    for name in ("diagonal", "dense", "low_rank"):
        case = run[name]
        for key, want in case.items():
            if key == "A":
                continue
            A = case["A"].clone()
            if key[0] == "symeig_psd":
                got = eig.symeig_psd(A, eigenvectors=True, upper=key[1], shift=key[2])
            elif key[0] == "symeig":
                got = eig.symeig(A, eigenvectors=True, upper=key[1])
            else:
                got = eig.symeig(A, eigenvectors=False, upper=key[1])
                assert got[1].numel() == 0
            _same_pairs(got, want, f"{name} {key}")
            assert torch.equal(A, case["A"])  # the input is left alone
    rect = run["shift_diag_rectangular"]
    assert torch.equal(eig.shift_diag(rect["input"], rect["shift"]), rect["result"])
    assert eig.shift_diag(rect["input"], 0.0) is rect["input"]


def test_eig_utils_contract(monkeypatch):
    """``test/utils/test_stable_symeig.py:49-76,102-122`` and the error contract of ``utils/eig.py``."""
    from vivit_b200.utils import eig

    double.install(monkeypatch)
    A = RUN["__eig_utils__"]["low_rank"]["A"]
    backup = A.clone()
    eig.symeig_psd(A, shift=1.0, shift_inplace=False)
    assert torch.equal(A, backup)
    eig.symeig_psd(A, shift=1.0, shift_inplace=True)  # shifted and shifted back: equal up to rounding
    assert torch.allclose(A, backup, atol=1e-12)
    evals, evecs = eig.symeig(backup, eigenvectors=True)
    assert evals.numel() == 5 and evecs.shape == (12, 5)  # rank 5: seven numerically-zero pairs dropped
    with pytest.raises(ValueError):
        eig.symeig_psd(torch.zeros(2, 2, 2))
    with pytest.raises(ValueError):
        eig.symeig(torch.zeros(4))
    with pytest.raises(RuntimeError, match="NaN"):
        eig.symeig(torch.full((2, 2), float("nan"), dtype=torch.float64))


# ---- vivit/utils/gram.py and vivit/utils/ggn.py ------------------------------------


def test_gram_and_ggn_utils_reproduce_the_reference_run(monkeypatch):
    import types

    from vivit_b200.utils import ggn, gram

    double.install(monkeypatch)
    want = RUN["__gram_utils__"]
    params = [types.SimpleNamespace(shape=s, vt=vt, gb=gb) for s, vt, gb in zip(want["shapes"], want["vt"], want["gb"])]
    mat_cn, mat_sqrt = want["mat_cn"], want["mat_sqrt"]

    def same(got, key):
        w = want[key]
        if isinstance(w, (list, tuple)):
            assert len(got) == len(w)
            for g, ww in zip(got, w):
                close(g, ww, what=key)
        else:
            close(got, w, what=key)

    same(gram.pairwise_dot(params[0].gb, start_dim=1), "pairwise_dot_1")
    same(gram.pairwise_dot(params[0].vt, start_dim=2), "pairwise_dot_2")
    same(gram.pairwise_dot(params[0].vt, start_dim=2, flatten=False), "pairwise_dot_2_unflattened")
    same(gram.partial_contract(params[0].vt, want["other"], (2, 1)), "partial_contract")
    same(gram.reshape_as_square(gram.pairwise_dot(params[1].vt, 2, flatten=False)), "reshape_as_square")
    same(gram.compute_gram_mat(params, "vt", 2), "compute_gram_mat")
    same(gram.compute_gram_mat(params, "vt", 2, flatten=False), "compute_gram_mat_unflattened")
    same(gram.compute_gram_mat(params, "gb", 1), "compute_gram_mat_grad")
    same(gram.sqrt_gram_mat_prod(mat_sqrt, params, "vt", 2), "sqrt_gram_mat_prod")
    same(gram.sqrt_gram_mat_prod(mat_sqrt, params, "vt", 2, concat=True), "sqrt_gram_mat_prod_concat")
    same([gram.mVp(p.vt, m, 2) for p, m in zip(params, want["mats_p"])], "mVp")
    same([ggn.Vmp(p.vt, mat_cn, 2) for p in params], "Vmp")
    same(ggn.V_mat_prod(mat_cn, params, "vt"), "V_mat_prod")
    same(ggn.V_mat_prod(mat_cn, params, "vt", concat=True), "V_mat_prod_concat")
    same(ggn.V_mat_prod(mat_cn[:, :, :2].contiguous(), params, "vt", subsampling=[3, 1]), "V_mat_prod_sub")

    with pytest.raises(ValueError):
        gram.partial_contract(params[0].vt, want["other"], (2, 2))
    with pytest.raises(NotImplementedError):
        gram.sqrt_gram_mat_prod(mat_cn, params, "vt", 2)
    with pytest.raises(ValueError):
        list(gram.split_list([1, 2, 3], [1, 1]))
    assert list(gram.split_list("abcde", [2, 3])) == ["ab", "cde"]
    assert gram.compute_gram_mat([], "vt", 2) is None
