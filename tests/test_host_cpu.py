"""Host-side logic on CPU: backprop engine, factors, parameter-group hooks, the four
Computations and their error contract.

There is no GPU here and the product has no CPU path, so the kernel layer is replaced by
the plain-torch test double (``tests/_torch_kernels.py``, installed with ``monkeypatch``);
everything above it is the shipped code.  Results are compared with the oracle restatement
of the reference on the same seeded inputs (float64) and with the autograd ground truth.
"""

import warnings

import pytest
import torch
from torch import nn

import tests._torch_kernels as double
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


@pytest.fixture(autouse=True)
def torch_kernels(monkeypatch):
    double.install(monkeypatch)


def run_backward(model, loss_fn, x, y, exts, hook):
    from vivit_b200 import backpack, extend

    model, loss_fn = extend(model), extend(loss_fn)
    loss = loss_fn(model(x), y)
    with backpack(*exts, extension_hook=hook):
        loss.backward()
    for p in model.parameters():
        p.grad = None


def close(a, b, rtol=1e-8, atol=1e-11):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a - b).abs().max()


SUBS = [None, [1, 0]]


@pytest.mark.parametrize("grouping", GROUPINGS, ids=GROUPING_IDS)
@pytest.mark.parametrize("sub", SUBS, ids=["full", "sub10"])
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_eigvalsh_matches_oracle(problem, sub, grouping):
    from vivit_b200 import EigvalshComputation

    model, loss, x, y = problem.make(torch.float64)
    groups = grouping(model)
    comp = EigvalshComputation(subsampling=sub)
    run_backward(model, loss, x, y, [comp.get_extension()], comp.get_extension_hook(groups))
    want = ref.eigvalsh(model, loss, x, y, groups, subsampling=sub)
    for g, w in zip(groups, want):
        close(comp.get_result(g), w)


@pytest.mark.parametrize("grouping", GROUPINGS, ids=GROUPING_IDS)
@pytest.mark.parametrize("sub", SUBS, ids=["full", "sub10"])
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_eigh_matches_oracle_and_ground_truth(problem, sub, grouping):
    from vivit_b200 import EighComputation

    model, loss, x, y = problem.make(torch.float64)
    groups = grouping(model, criterion=keep_nonzero)
    comp = EighComputation(subsampling=sub)
    run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))
    want = ref.eigh(model, loss, x, y, groups, subsampling=sub)
    truth = AutogradGGN(model, loss, x, y)
    for g, (w_evals, w_evecs) in zip(groups, want):
        evals, evecs = comp.get_result(g)
        close(evals, w_evals)
        assert [e.shape for e in evecs] == [(evals.numel(), *p.shape) for p in g["params"]]
        # eigenvectors up to sign: compare projectors  e e^T
        flat = torch.cat([e.flatten(1) for e in evecs], 1)
        wflat = torch.cat([e.flatten(1) for e in w_evecs], 1)
        close(flat.t() @ flat, wflat.t() @ wflat, 1e-6, 1e-9)
        Ge = truth.ggn_mat_prod(g["params"], evecs, sub)
        for a, e in zip(Ge, evecs):
            close(a, torch.einsum("i,i...->i...", evals, e), 1e-6, 1e-9)
    for p in model.parameters():
        assert not hasattr(p, "vivit_ggn_exact")  # savefields are freed (eigh.py:270)


@pytest.mark.parametrize("k", [1, 10])
@pytest.mark.parametrize("sub_ggn", [None, [0, 1]], ids=["ggn-full", "ggn-01"])
@pytest.mark.parametrize("sub_grad", [None, [0, 1]], ids=["grad-full", "grad-01"])
@pytest.mark.parametrize("grouping", GROUPINGS, ids=GROUPING_IDS)
@pytest.mark.parametrize("problem", PROBLEMS, ids=IDS)
def test_optim_computations_match_oracle(problem, grouping, sub_grad, sub_ggn, k):
    from vivit_b200 import DirectionalDampedNewtonComputation, DirectionalDerivativesComputation

    model, loss, x, y = problem.make(torch.float64)
    groups = grouping(model, criterion=make_top_k(k), damping=constant_damping(1.0))
    comp = DirectionalDerivativesComputation(subsampling_grad=sub_grad, subsampling_ggn=sub_ggn)
    run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))
    want = ref.directional_derivatives(model, loss, x, y, groups, sub_grad, sub_ggn)
    for g, (wg, wl) in zip(groups, want):
        gam, lam = comp.get_result(g)
        close(gam.abs(), wg.abs(), 1e-7, 1e-10)
        close(lam, wl, 1e-7, 1e-10)
    newton = DirectionalDampedNewtonComputation(subsampling_grad=sub_grad, subsampling_ggn=sub_ggn)
    run_backward(model, loss, x, y, newton.get_extensions(), newton.get_extension_hook(groups))
    want = ref.directional_damped_newton(model, loss, x, y, groups, sub_grad, sub_ggn)
    for g, w in zip(groups, want):
        steps = newton.get_result(g)
        assert [s.shape for s in steps] == [p.shape for p in g["params"]]
        for s, t in zip(steps, w):
            close(s, t, 1e-7, 1e-10)
    for p in model.parameters():
        assert not hasattr(p, "sqrt_ggn_exact") and not hasattr(p, "grad_batch")


@pytest.mark.parametrize("problem", PROBLEMS[:2], ids=IDS[:2])
def test_mc_computations_on_pinned_samples(problem):
    """MC factor with the class ids pinned on both sides."""
    from vivit_b200 import DirectionalDampedNewtonComputation, EighComputation

    model, loss, x, y = problem.make(torch.float64)
    sub = [0, 2]
    with torch.no_grad():
        ids = ref.sample_ce_classes(model(x), sub, 1)
    groups = [{"params": list(model.parameters()), "criterion": make_top_k(2), "damping": constant_damping(1.0)}]
    comp = EighComputation(subsampling=sub, mc_samples=1)
    comp._mc_state = ids
    run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))
    (w_evals, _), = ref.eigh(model, loss, x, y, groups, subsampling=sub, mc_samples=1, mc_state=ids)
    close(comp.get_result(groups[0])[0], w_evals)
    newton = DirectionalDampedNewtonComputation(subsampling_ggn=sub, mc_samples_ggn=1)
    newton._mc_state = ids
    run_backward(model, loss, x, y, newton.get_extensions(), newton.get_extension_hook(groups))
    (want,) = ref.directional_damped_newton(model, loss, x, y, groups, None, sub, 1, ids)
    for s, t in zip(newton.get_result(groups[0]), want):
        close(s, t, 1e-7, 1e-10)


@pytest.mark.parametrize("problem", PROBLEMS + [PROBLEM_SUM], ids=IDS + ["mlp-ce-sum"])
def test_savefield_closures_and_materialised_extensions(problem):
    """Lower-level functional API and the [BackPACK]-shaped tensors."""
    from vivit_b200 import BatchGrad, SqrtGGNExact, ViViTGGNExact

    model, loss, x, y = problem.make(torch.float64)
    sub = [0, 0, 1] if x.shape[0] >= 2 else None
    run_backward(model, loss, x, y, [ViViTGGNExact(subsampling=sub), SqrtGGNExact(subsampling=sub), BatchGrad()], None)
    sweep = ref.backward_sweep(model, loss, x, y, subsampling_ggn=sub, want_vivit=True, want_sqrt_ggn=True, want_grad_batch=True)
    torch.manual_seed(3)
    for p in model.parameters():
        close(p.sqrt_ggn_exact, sweep.sqrt_ggn[id(p)])
        close(p.grad_batch, sweep.grad_batch[id(p)])
        cl, want = p.vivit_ggn_exact, sweep.vivit[id(p)]
        close(cl["gram_mat"](), want["gram_mat"]())
        C, N = p.sqrt_ggn_exact.shape[:2]
        mat = torch.rand(3, C, N, dtype=torch.float64)
        close(cl["V_mat_prod"](mat), want["V_mat_prod"](mat))
        pm = torch.rand(4, *p.shape, dtype=torch.float64)
        close(cl["V_t_mat_prod"](pm), want["V_t_mat_prod"](pm))


def test_hook_fires_group_when_last_param_arrives():
    """Per-layer groups finish layer by layer, a full-network group only at the first layer
    (vivit/utils/hooks.py:309-330)."""
    from vivit_b200 import EigvalshComputation, backpack, extend

    torch.manual_seed(0)
    model = extend(nn.Sequential(nn.Linear(5, 4), nn.ReLU(), nn.Linear(4, 3)).double())
    loss_fn = extend(nn.CrossEntropyLoss())
    x, y = torch.rand(4, 5, dtype=torch.float64), torch.randint(0, 3, (4,))
    layers = [model[0], model[2]]
    groups = [{"params": list(l.parameters())} for l in layers]
    comp = EigvalshComputation(batch_solves=False)  # the reference's order: decompose inside every group's hook
    hook = comp.get_extension_hook(groups)
    seen = []

    def spy(module):
        hook(module)
        seen.append((type(module).__name__, [id(g) in comp._evals for g in groups]))

    with backpack(comp.get_extension(), extension_hook=spy):
        loss_fn(model(x), y).backward()
    names = [n for n, _ in seen]
    assert names == ["CrossEntropyLoss", "Linear", "ReLU", "Linear"]
    assert seen[1][1] == [False, True] and seen[3][1] == [True, True]


def test_error_contract():
    from vivit_b200 import (
        DirectionalDampedNewtonComputation,
        DirectionalDerivativesComputation,
        EighComputation,
        EigvalshComputation,
        ViViTGGNExact,
        backpack,
        extend,
    )

    for cls in (EigvalshComputation, EighComputation, DirectionalDerivativesComputation, DirectionalDampedNewtonComputation):
        with pytest.raises(KeyError):
            cls().get_result({"params": []})  # test_eigh.py:179-184 et al.
    with pytest.raises(ValueError):
        EighComputation(subsampling=[0, 0])
    with pytest.raises(ValueError):
        DirectionalDerivativesComputation(subsampling_ggn=[1, 1])
    with pytest.raises(AssertionError):
        DirectionalDampedNewtonComputation(mc_samples_ggn=2)
    p = nn.Parameter(torch.zeros(2))
    with pytest.raises(ValueError):
        EighComputation().get_extension_hook([{"params": [p]}])  # no criterion
    with pytest.raises(ValueError):
        EigvalshComputation().get_extension_hook([{"criterion": keep_all}])  # no params
    with pytest.raises(ValueError):
        EigvalshComputation().get_extension_hook([{"params": [p]}, {"params": [p]}])
    with pytest.raises(ValueError):
        DirectionalDampedNewtonComputation().get_extension_hook([{"params": [p], "criterion": keep_all}])
    with pytest.raises(ValueError):
        backpack(ViViTGGNExact)  # class instead of instance
    model = extend(nn.Sequential(nn.Linear(3, 2), nn.Softplus()))
    loss_fn = extend(nn.CrossEntropyLoss())
    with pytest.raises(NotImplementedError):
        with backpack(ViViTGGNExact()):
            loss_fn(model(torch.rand(2, 3)), torch.tensor([0, 1])).backward()


def test_small_eigenvalue_warning():
    from vivit_b200 import EighComputation

    model, loss, x, y = PROBLEMS[0].make(torch.float64)
    groups = [{"params": list(model.parameters()), "criterion": keep_all}]
    comp = EighComputation()
    with pytest.warns(UserWarning, match="small eigenvalues"):
        run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))
    comp = EighComputation(warn_small_eigvals=0)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        run_backward(model, loss, x, y, comp.get_extensions(), comp.get_extension_hook(groups))


def test_no_context_no_side_effects():
    from vivit_b200 import extend

    model = extend(nn.Linear(3, 2))
    model(torch.rand(2, 3)).sum().backward()
    assert not hasattr(model.weight, "vivit_ggn_exact")


def test_criterion_results_follow_the_reference_indexing():
    """``evals[keep]`` semantics of the reference (``vivit/linalg/eigh.py:252-265``): positions, negative
    positions, boolean masks (a mask used to be cast to 0/1 positions)."""
    from vivit_b200.utils import keep_indices

    ev = torch.tensor([0.0, 1e-9, 0.5, 2.0, 3.0])
    assert keep_indices([2, 4], ev).tolist() == [2, 4]
    assert keep_indices([-1], ev).tolist() == [4]
    assert keep_indices(torch.tensor([-2, 0]), ev).tolist() == [3, 0]
    assert keep_indices(ev > 1e-6, ev).tolist() == [2, 3, 4]
    assert keep_indices([True, False, False, False, True], ev).tolist() == [0, 4]
    assert keep_indices([], ev).tolist() == []
    for bad in ([5], [-6], [0.5], torch.tensor([True, False])):
        with pytest.raises(IndexError):
            keep_indices(bad, ev)


def test_boolean_mask_criterion_end_to_end():
    """A criterion returning ``evals > tol`` keeps exactly the eigenpairs the same criterion written with
    positions keeps."""
    import vivit_b200 as vv

    results = []
    for crit in (lambda ev: ev > 1e-6, lambda ev: [i for i in range(ev.numel()) if ev[i] > 1e-6]):
        model, loss_fn, x, y = PROBLEMS[0].make()
        groups = [{"params": list(model.parameters()), "criterion": crit}]
        comp = vv.EighComputation()
        run_backward(model, loss_fn, x, y, [comp.get_extension()], comp.get_extension_hook(groups))
        results.append(comp.get_result(groups[0]))
    assert results[0][0].numel() == results[1][0].numel() > 0
    assert torch.allclose(results[0][0], results[1][0])


def test_step_leaves_no_garbage():
    """One backward pass per Computation leaves nothing for Python's cyclic collector: every factor, Gram matrix
    and result is released by its reference count the moment the caller drops it.  (A cycle through the hook's
    closures or through the weight/bias factor links of a Linear layer would hold gigabytes of factor tensors at
    bench size until a generation-2 collection -- the 60 ... 200 ms outlier steps of round 2.)"""
    import gc

    import vivit_b200 as vv

    model, loss_fn, x, y = PROBLEMS[0].make()
    groups = [
        {"params": list(model.parameters()), "criterion": keep_nonzero, "damping": constant_damping(1.0)}
    ]

    def step(cls):
        comp = cls()
        exts = comp.get_extensions() if hasattr(comp, "get_extensions") else [comp.get_extension()]
        run_backward(model, loss_fn, x, y, exts, comp.get_extension_hook(groups))
        comp.get_result(groups[0])

    classes = (
        vv.EighComputation,
        vv.EigvalshComputation,
        vv.DirectionalDerivativesComputation,
        vv.DirectionalDampedNewtonComputation,
    )
    for cls in classes:  # first calls: caches, lazy imports
        step(cls)
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    saved = gc.get_debug()
    gc.set_debug(gc.DEBUG_SAVEALL)
    try:
        for cls in classes:
            step(cls)
        gc.collect()
        tensors = [o for o in gc.garbage if isinstance(o, torch.Tensor)]
        kinds = sorted({type(o).__name__ for o in gc.garbage})
    finally:
        gc.set_debug(saved)
        gc.garbage.clear()
        if was_enabled:
            gc.enable()
    assert not tensors, (len(tensors), kinds)


@pytest.mark.parametrize("grouping", GROUPINGS, ids=GROUPING_IDS)
def test_shared_solve_queue_gives_the_immediate_results(grouping, monkeypatch):
    """``EighComputation``, ``DirectionalDerivativesComputation`` and ``DirectionalDampedNewtonComputation`` on ONE
    ``SolveQueue``: three backward passes,
    nothing decomposed until the first ``get_result``, then one batched call per matrix shape -- and the results
    of the immediate order (``eigh.py:248``, ``directional_derivatives.py:291``)."""
    import vivit_b200 as vv
    from vivit_b200 import kernels

    problem = PROBLEMS[0]
    calls = []
    single, batched = kernels.syevj, kernels.syevj_batched
    monkeypatch.setattr(kernels, "syevj", lambda G, *a, **k: (calls.append(1), single(G, *a, **k))[1])
    monkeypatch.setattr(kernels, "syevj_batched", lambda G, *a, **k: (calls.append(G.shape[0]), batched(G, *a, **k))[1])

    def both(queue):
        model, loss_fn, x, y = problem.make()
        groups = grouping(model, criterion=keep_nonzero, damping=constant_damping(1.0))
        kw = {} if queue is None else {"solve_queue": queue}
        eigh, dirs = vv.EighComputation(**kw), vv.DirectionalDerivativesComputation(**kw)
        newton = vv.DirectionalDampedNewtonComputation(**kw)
        run_backward(model, loss_fn, x, y, [eigh.get_extension()], eigh.get_extension_hook(groups))
        run_backward(model, loss_fn, x, y, dirs.get_extensions(), dirs.get_extension_hook(groups))
        run_backward(model, loss_fn, x, y, newton.get_extensions(), newton.get_extension_hook(groups))
        if queue is not None:
            assert len(queue) == 3 * len(groups) and not calls
        return [(eigh.get_result(g), dirs.get_result(g), newton.get_result(g)) for g in groups]

    want = both(None)
    del calls[:]
    got = both(vv.SolveQueue())
    assert sum(calls) == 3 * len(want) and len(calls) < sum(calls)  # batched
    for ((ev0, vecs0), (g0, l0), s0), ((ev1, vecs1), (g1, l1), s1) in zip(want, got):
        close(ev0, ev1)
        close(g0, g1)
        close(l0, l1)
        for a, b in zip(list(vecs0) + list(s0), list(vecs1) + list(s1)):
            close(a, b)


def test_solve_queue_keeps_the_other_results_when_a_callback_raises():
    """A criterion that raises inside the deferred half of one Computation's hook is reported once, by the
    ``get_result`` that flushed the queue, and does not cost the other Computation its results."""
    import vivit_b200 as vv

    def broken(evals):
        raise RuntimeError("criterion failed")

    model, loss_fn, x, y = PROBLEMS[0].make()
    good = [{"params": list(model.parameters()), "criterion": keep_nonzero}]
    bad = [{"params": list(model.parameters()), "criterion": broken}]
    queue = vv.SolveQueue()
    eigh, dirs = vv.EighComputation(solve_queue=queue), vv.DirectionalDerivativesComputation(solve_queue=queue)
    run_backward(model, loss_fn, x, y, [eigh.get_extension()], eigh.get_extension_hook(bad))
    run_backward(model, loss_fn, x, y, dirs.get_extensions(), dirs.get_extension_hook(good))
    with pytest.raises(RuntimeError, match="criterion failed"):
        eigh.get_result(bad[0])
    assert len(queue) == 0
    gammas, lambdas = dirs.get_result(good[0])
    assert gammas.shape[1] == lambdas.shape[1] > 0
    with pytest.raises(KeyError):
        eigh.get_result(bad[0])


def test_solve_queue_buckets_by_shape_and_by_vectors():
    """Matrices of different sizes, and eigenvalue-only requests (``EigvalshComputation``), share a queue but not a
    batched call; every Computation gets the results of its immediate order."""
    import vivit_b200 as vv
    from vivit_b200 import kernels

    queue = vv.SolveQueue()
    done = {}
    g = torch.Generator().manual_seed(0)
    mats = []
    for i, n in enumerate((5, 7, 5, 5)):
        a = torch.randn(n, n + 2, generator=g, dtype=torch.float64)
        mats.append(a @ a.t())
    for i, m in enumerate(mats):
        queue.submit(m, lambda ev, vec, i=i: done.__setitem__(i, (ev, vec)), vectors=i != 3)
    assert len(queue) == 4
    queue.flush()
    assert len(queue) == 0 and sorted(done) == [0, 1, 2, 3]
    for i, m in enumerate(mats):
        ev, vec = done[i]
        close(ev, torch.linalg.eigvalsh(m), rtol=1e-8, atol=1e-10)
        assert (vec is None) == (i == 3)

    model, loss_fn, x, y = PROBLEMS[0].make()
    groups = GROUPINGS[1](model, criterion=keep_nonzero)
    plain, queued = vv.EigvalshComputation(), vv.EigvalshComputation(solve_queue=queue)
    run_backward(model, loss_fn, x, y, [plain.get_extension()], plain.get_extension_hook(groups))
    run_backward(model, loss_fn, x, y, [queued.get_extension()], queued.get_extension_hook(groups))
    assert len(queue) == len(groups)
    for grp in groups:
        close(queued.get_result(grp), plain.get_result(grp))


def test_pool_output_size_is_torch_s():
    """``kernels.pool_output_size`` (host arithmetic of the native arg-max kernel) against the shapes torch's pooling
    produces, floor and ceil mode."""
    import torch.nn.functional as F

    from vivit_b200.kernels import pool_output_size

    for size in (5, 8, 9, 11, 32):
        for k, s, p, d in [(2, 2, 0, 1), (3, 2, 0, 1), (3, 2, 1, 1), (3, 1, 1, 1), (3, 3, 0, 1), (2, 1, 0, 2), (4, 3, 1, 1),
                           (3, 2, 1, 2), (5, 4, 2, 1)]:
            if d * (k - 1) + 1 > size + 2 * p:
                continue
            for ceil in (False, True):
                want = F.max_pool2d(torch.zeros(1, 1, size, size), k, s, p, d, ceil).shape[-1]
                assert pool_output_size(size, k, s, p, d, ceil) == want, (size, k, s, p, d, ceil)
