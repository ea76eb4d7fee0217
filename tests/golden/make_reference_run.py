"""Golden vectors produced by the REFERENCE'S OWN CODE, executed in the build container.

    python tests/golden/make_reference_run.py        # writes tests/golden/reference_run.pt

``import vivit`` needs BackPACK (absent here) and ``Tensor.symeig`` (removed in
torch 2).  What BackPACK contributes to this path are per-parameter TENSORS -- ``grad_batch``,
``sqrt_ggn_exact`` -- which the reference's hooks then consume; everything after that (Gram assembly,
rescaling, ``symeig``, filtering, gammas, lambdas, Newton coefficients, back-transformation, normalisation,
the hook and group bookkeeping) is the reference's own code in ``vivit/{linalg,optim,utils}``.  So:

* ``backpack.*`` is replaced by import stubs that carry NAMES only (a meta-path finder that fabricates
  empty classes; ``savefield`` strings and the ``LossHessianStrategy`` constants as published);
* ``Tensor.symeig(eigenvectors, upper)`` is shimmed to ``torch.linalg.eigh(UPLO=...)`` -- the replacement
  torch's deprecation notice prescribed (same order, same column convention);
* the per-parameter tensors are computed by plain autograd (per-sample output Jacobians; the loss
  Hessian of each sample by double backward, factored symmetrically by ``eigh``), independently of
  ``oracle/reference_path.py``;
* for the two ``linalg`` classes the savefield is the dict of closures of ``base.py:94-130``, built here
  from the reference's own ``vivit.utils.gram.pairwise_dot`` / ``mVp`` and ``vivit.utils.ggn.Vmp``;
* then the reference's ``EigvalshComputation``, ``EighComputation``, ``DirectionalDerivativesComputation``
  and ``DirectionalDampedNewtonComputation`` run unmodified: ``get_extension_hook(param_groups)`` is called
  on every leaf module in reverse order, as BackPACK would, and ``get_result(group)`` is stored;
* the extension hooks of ``vivit.extensions.hooks`` (``GramBatchGrad``, ``CenteredGramBatchGrad``,
  ``CenteredBatchGrad``, ``GramSqrtGGNExact``) run the same way on the full batch (``"__gram_hooks__"``);
* ``ViViTGGNLinear.weight`` (``linear.py:29-81``), the structured closures of a Linear weight, runs on
  seeded tensors (``"__linear_closures__"``; ``backpack.utils.subsampling.subsample`` is the one helper
  given a body: keep the listed samples);
* ``vivit/utils/eig.py`` (``symeig_psd``, ``symeig`` with its zero-eigenvalue filter, ``shift_diag``) runs on the
  matrices of the reference's ``test_stable_symeig.py`` and on a rank-deficient Gram matrix (``"__eig_utils__"``);
* ``vivit/utils/gram.py`` and ``vivit/utils/ggn.py`` (``pairwise_dot``, ``partial_contract``, ``compute_gram_mat``,
  ``sqrt_gram_mat_prod``, ``mVp``, ``Vmp``, ``V_mat_prod`` ...) run on seeded tensors (``"__gram_utils__"``);
* ``vivit/hessianfree/lanczos.py`` and ``utils.py`` need no BackPACK: they run on a seeded dense symmetric
  matrix with ``numpy.random.seed`` fixed and explicit spectrum boundaries (``"__lanczos__"``).

Any symmetric factor of the loss Hessian gives the same GGN, hence the same eigenvalues, directional
derivatives and Newton steps; eigenvectors are compared through projectors.  ``/root/reference`` does
not exist on the GPU box: only the committed ``.pt`` file travels.
"""

import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)

from tests.problems import GROUPING_IDS, GROUPINGS, PROBLEMS, constant_damping, keep_nonzero, make_top_k  # noqa: E402

SAVEFIELDS = {"BatchGrad": "grad_batch", "SqrtGGNExact": "sqrt_ggn_exact", "SqrtGGNMC": "sqrt_ggn_mc"}


# ---- import stubs for backpack.* (names only) ------------------------------------------------------
class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _StubMeta(name, (_Stub,), {})


class _Stub(metaclass=_StubMeta):
    savefield = None

    def __init__(self, *args, savefield=None, subsampling=None, **kwargs):
        if savefield is not None:
            self.savefield = savefield
        self._subsampling = subsampling

    def get_subsampling(self):
        return self._subsampling


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        attrs = {"savefield": SAVEFIELDS.get(name)}
        if name == "LossHessianStrategy":  # published string constants of backpack.extensions.secondorder.hbp
            attrs.update(EXACT="exact", SAMPLING="sampling", SUMMARY="summary")
        cls = _StubMeta(name, (_Stub,), attrs)
        setattr(self, name, cls)
        return cls


def _subsample(tensor, dim=0, subsampling=None):
    """``backpack.utils.subsampling.subsample`` as published: keep the listed entries along ``dim``."""
    if subsampling is None:
        return tensor
    return tensor.index_select(dim, torch.tensor(subsampling, device=tensor.device))


class _BackpackFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname == "backpack" or fullname.startswith("backpack."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        module = _StubModule(spec.name)
        module.__path__ = []
        if spec.name == "backpack.utils.subsampling":
            module.subsample = _subsample  # the one BackPACK helper the structured Linear path calls
        return module

    def exec_module(self, module):
        pass


def _symeig(self, eigenvectors=False, upper=True):
    uplo = "U" if upper else "L"
    if eigenvectors:
        return torch.linalg.eigh(self, UPLO=uplo)
    return torch.linalg.eigvalsh(self, UPLO=uplo), torch.empty(0, dtype=self.dtype, device=self.device)


def import_reference():
    sys.meta_path.insert(0, _BackpackFinder())
    torch.Tensor.symeig = _symeig  # torch 2 keeps the name only to raise the deprecation error
    sys.path.insert(0, REFERENCE)
    import vivit

    assert os.path.realpath(vivit.__file__).startswith(REFERENCE), vivit.__file__
    return vivit


# ---- what BackPACK would have left on the parameters, by autograd -----------------------------------
def per_parameter_tensors(model, loss_fn, x, y):
    """``sqrt_ggn_exact`` ``[F, N, *p.shape]`` and ``grad_batch`` ``[N, *p.shape]`` per parameter."""
    params = [p for p in model.parameters() if p.requires_grad]
    out = model(x)
    N = out.shape[0]
    F = out[0].numel()
    flat = out.reshape(N, F)
    jac = [torch.zeros(N, F, *p.shape, dtype=out.dtype) for p in params]  # d out[n, f] / d p
    for n in range(N):
        for f in range(F):
            grads = torch.autograd.grad(flat[n, f], params, retain_graph=True, allow_unused=True)
            for j, g in zip(jac, grads):
                if g is not None:
                    j[n, f] = g
    o = out.detach().clone().requires_grad_(True)
    loss = loss_fn(o, y)
    (g_out,) = torch.autograd.grad(loss, o, create_graph=True)  # includes the 1/N of the mean
    g_flat = g_out.reshape(N, F)
    factors = []
    for n in range(N):
        rows = []
        for f in range(F):
            (h,) = torch.autograd.grad(g_flat[n, f], o, retain_graph=True)
            rows.append(h.reshape(N, F)[n])
        H = torch.stack(rows)
        H = 0.5 * (H + H.t())
        lam, Q = torch.linalg.eigh(H)
        S = Q * lam.clamp(min=0.0).sqrt()  # S S^T = H_n (PSD for CE / MSE)
        assert torch.allclose(S @ S.t(), H, atol=1e-12), "loss Hessian is not PSD?"
        factors.append(S)
    S = torch.stack(factors)  # [N, F(out), F(factor column)]
    sqrt_ggn = [torch.einsum("nov,no...->vn...", S, j) for j in jac]
    grad_batch = [torch.einsum("no,no...->n...", g_flat.detach(), j) for j in jac]
    return params, sqrt_ggn, grad_batch


def leaf_modules_reversed(model):
    leaves = [m for m in model.modules() if next(m.children(), None) is None]
    return leaves[::-1]


def run_hook(model, x, hook):
    for module in leaf_modules_reversed(model):
        module.input0 = x  # only its batch size is read (linalg/utils.py:31-64)
        hook(module)
        del module.input0


def gram_hooks(vivit):
    """``vivit.extensions.hooks`` on the full batch of every problem (gram_batch_grad.py:7-213,
    gram_sqrt_ggn.py:9-142)."""
    from vivit.extensions import hooks as ref_hooks

    out = {}
    for problem in PROBLEMS:
        model, loss_fn, x, y = problem.make(torch.float64)
        params, sqrt_ggn, grad_batch = per_parameter_tensors(model, loss_fn, x, y)
        case = {}
        for name, cls in (("gram", ref_hooks.GramBatchGrad), ("centered_gram", ref_hooks.CenteredGramBatchGrad)):
            for p, g in zip(params, grad_batch):
                p.grad_batch = g.clone()
            hook = cls(layerwise=True, free_grad_batch=True)
            for module in leaf_modules_reversed(model):
                hook(module)
            case[name] = hook.get_result().clone()
            case[name + "_layerwise"] = [getattr(p, hook.savefield).clone() for p in params]
            for p in params:
                assert not hasattr(p, "grad_batch")
                delattr(p, hook.savefield)
        for p, g in zip(params, grad_batch):
            p.grad_batch = g.clone()
        hook = ref_hooks.CenteredBatchGrad()
        for module in leaf_modules_reversed(model):
            hook(module)
        case["centered_grad_batch"] = [getattr(p, hook.savefield).clone() for p in params]
        for p, V in zip(params, sqrt_ggn):
            delattr(p, hook.savefield)
            del p.grad_batch
            p.sqrt_ggn_exact = V.clone()
        hook = ref_hooks.GramSqrtGGNExact(free_sqrt_ggn=True)
        for module in leaf_modules_reversed(model):
            hook(module)
        # the Gram matrix depends on the loss-Hessian factor through an orthogonal change of basis per
        # sample: its spectrum does not
        case["gram_sqrt_ggn_evals"] = torch.linalg.eigvalsh(hook.get_result())
        out[problem.name] = case
    return out


def linear_closures():
    """The structured closures of ``ViViTGGNLinear.weight`` (``linear.py:29-81``) on seeded tensors, with and
    without sub-sampling.  ``derivatives._get_additional_dims`` (BackPACK) is told there are none."""
    from vivit.extensions.secondorder.vivit.linear import ViViTGGNLinear

    ext_module = ViViTGGNLinear()
    ext_module.derivatives = types.SimpleNamespace(_get_additional_dims=lambda module: ())
    out = []
    gen = torch.Generator().manual_seed(0)
    for C, N, n_out, n_in, sub in ((3, 4, 5, 6, None), (2, 5, 3, 4, [2, 0, 4]), (10, 7, 10, 33, None)):
        n_sub = N if sub is None else len(sub)
        s = torch.randn(C, n_sub, n_out, generator=gen, dtype=torch.float64)
        x = torch.randn(N, n_in, generator=gen, dtype=torch.float64)
        mat_v = torch.randn(2, C, n_sub, generator=gen, dtype=torch.float64)
        mat_vt = torch.randn(2, n_out, n_in, generator=gen, dtype=torch.float64)
        module = types.SimpleNamespace(input0=x)
        ext = types.SimpleNamespace(get_subsampling=lambda sub=sub: sub)
        fns = ext_module.weight(ext, module, None, None, s)
        out.append({
            "s": s, "input0": x, "subsampling": sub, "mat_v": mat_v, "mat_vt": mat_vt,
            "gram_mat": fns["gram_mat"](), "V_mat_prod": fns["V_mat_prod"](mat_v),
            "V_t_mat_prod": fns["V_t_mat_prod"](mat_vt),
        })
    return out


def eig_utils():
    """``vivit/utils/eig.py:6-134`` (``symeig_psd``, ``symeig``, ``remove_zero_evals``, ``shift_diag``) on the
    matrices of ``test/utils/test_stable_symeig.py:9-24`` and on a seeded rank-deficient Gram matrix."""
    from vivit.utils import eig as ref_eig

    gen = torch.Generator().manual_seed(0)
    B = torch.randn(12, 5, generator=gen, dtype=torch.float64)
    mats = {
        "diagonal": torch.diag(torch.tensor([1.1, 2.2, 9.9], dtype=torch.float64)),
        "dense": torch.tensor([[1.1, 2.2, 3.3], [4.4, 5.5, 6.6], [7.7, 8.8, 2.2]], dtype=torch.float64),
        "low_rank": B @ B.t(),
    }
    out = {}
    for name, A in mats.items():
        case = {"A": A}
        for upper in (True, False):
            for shift in (0.0, 0.1, 10.0):
                case[("symeig_psd", upper, shift)] = ref_eig.symeig_psd(
                    A.clone(), eigenvectors=True, upper=upper, shift=shift, shift_inplace=False
                )
            case[("symeig", upper)] = ref_eig.symeig(A.clone(), eigenvectors=True, upper=upper)
            case[("symeig_values_only", upper)] = ref_eig.symeig(A.clone(), eigenvectors=False, upper=upper)
        out[name] = case
    rect = torch.tensor([[1.0, 1.0], [2.0, 2.0], [3.0, 4.0]], dtype=torch.float64)
    out["shift_diag_rectangular"] = {"input": rect, "shift": 1.5, "result": ref_eig.shift_diag(rect, 1.5)}
    return out


def gram_utils():
    """``vivit/utils/gram.py`` and ``vivit/utils/ggn.py`` on seeded per-parameter tensors: two "parameters"
    of shapes ``[4, 3]`` and ``[5]`` carrying ``V^T [C=3, N=4, *shape]`` and per-sample gradients ``[N, *shape]``."""
    from vivit.utils import ggn as ref_ggn
    from vivit.utils import gram as ref_gram

    gen = torch.Generator().manual_seed(0)
    rand = lambda *shape: torch.randn(*shape, generator=gen, dtype=torch.float64)  # noqa: E731
    shapes = [(4, 3), (5,)]
    params = [types.SimpleNamespace(shape=s, vt=rand(3, 4, *s), gb=rand(4, *s)) for s in shapes]
    mat_cn, mat_sqrt = rand(2, 3, 4), rand(12, 6)
    mats_p = [rand(2, *s) for s in shapes]
    other = rand(7, 4, 3)
    out = {
        "shapes": shapes, "vt": [p.vt for p in params], "gb": [p.gb for p in params],
        "mat_cn": mat_cn, "mat_sqrt": mat_sqrt, "mats_p": mats_p, "other": other,
        "pairwise_dot_1": ref_gram.pairwise_dot(params[0].gb, start_dim=1),
        "pairwise_dot_2": ref_gram.pairwise_dot(params[0].vt, start_dim=2),
        "pairwise_dot_2_unflattened": ref_gram.pairwise_dot(params[0].vt, start_dim=2, flatten=False),
        "partial_contract": ref_gram.partial_contract(params[0].vt, other, (2, 1)),
        "reshape_as_square": ref_gram.reshape_as_square(ref_gram.pairwise_dot(params[1].vt, 2, flatten=False)),
        "compute_gram_mat": ref_gram.compute_gram_mat(params, "vt", 2),
        "compute_gram_mat_unflattened": ref_gram.compute_gram_mat(params, "vt", 2, flatten=False),
        "compute_gram_mat_grad": ref_gram.compute_gram_mat(params, "gb", 1),
        "sqrt_gram_mat_prod": ref_gram.sqrt_gram_mat_prod(mat_sqrt, params, "vt", 2),
        "sqrt_gram_mat_prod_concat": ref_gram.sqrt_gram_mat_prod(mat_sqrt, params, "vt", 2, concat=True),
        "mVp": [ref_gram.mVp(p.vt, m, 2) for p, m in zip(params, mats_p)],
        "Vmp": [ref_ggn.Vmp(p.vt, mat_cn, 2) for p in params],
        "V_mat_prod": ref_ggn.V_mat_prod(mat_cn, params, "vt"),
        "V_mat_prod_concat": ref_ggn.V_mat_prod(mat_cn, params, "vt", concat=True),
        "V_mat_prod_sub": ref_ggn.V_mat_prod(mat_cn[:, :, :2].contiguous(), params, "vt", subsampling=[3, 1]),
    }
    return out


def lanczos_vectors():
    """``vivit/hessianfree/lanczos.py:13-270`` and ``utils.py:7-57`` on a fixed symmetric matrix."""
    import importlib.util

    import numpy as np
    from scipy.sparse.linalg import aslinearoperator

    def load(name):
        spec = importlib.util.spec_from_file_location("ref_" + name, f"{REFERENCE}/vivit/hessianfree/{name}.py")
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        return module

    lanczos, utils = load("lanczos"), load("utils")
    rng = np.random.RandomState(0)
    B = rng.randn(48, 48)
    A = B @ B.T / 48 - 0.3 * np.eye(48)
    evals = np.linalg.eigvalsh(A)
    op = aslinearoperator(A)
    out = {"A": torch.from_numpy(A), "seed": 1}
    calls = {
        "fast_lanczos": dict(ncv=14),
        "fast_lanczos_tridiagonal": dict(ncv=14, use_eigh_tridiagonal=True),
        "lanczos_approximate_spectrum": dict(
            ncv=16, num_points=64, num_repeats=3, kappa=3.0, boundaries=(float(evals[0]), float(evals[-1]))
        ),
        "lanczos_approximate_log_spectrum": dict(
            ncv=16, num_points=64, num_repeats=3, kappa=1.04,
            boundaries=(float(np.abs(evals).min()), float(np.abs(evals).max())),
        ),
    }
    for name, kwargs in calls.items():
        np.random.seed(out["seed"])
        fn = getattr(lanczos, "fast_lanczos" if name.startswith("fast_lanczos") else name)
        res = fn(op, **kwargs)
        out[name] = {"kwargs": kwargs, "result": [torch.from_numpy(np.asarray(r, dtype=np.float64).copy()) for r in res]}
    c, V = rng.rand(5), np.linalg.qr(rng.randn(48, 5))[0]
    probe = rng.randn(48)
    out["low_rank"] = {
        "c": torch.from_numpy(c), "A": torch.from_numpy(V), "x": torch.from_numpy(probe),
        "LowRank": torch.from_numpy(utils.LowRank(c, V) @ probe),
        "Projector": torch.from_numpy(utils.Projector(V) @ probe),
    }
    return out


def main():
    vivit = import_reference()
    from vivit.utils.ggn import Vmp as ref_Vmp
    from vivit.utils.gram import mVp as ref_mVp
    from vivit.utils.gram import pairwise_dot as ref_pairwise_dot

    def closures(V_t):  # the savefield of ViViTGGNExact for a materialised factor (base.py:94-130)
        return {
            "gram_mat": lambda: ref_pairwise_dot(V_t, start_dim=2, flatten=False),
            "V_mat_prod": lambda mat: ref_Vmp(V_t, mat, 2),
            "V_t_mat_prod": lambda mat: ref_mVp(V_t, mat, 2),
        }

    out = {}
    for problem in PROBLEMS:
        model, loss_fn, x, y = problem.make(torch.float64)
        params, sqrt_ggn, grad_batch = per_parameter_tensors(model, loss_fn, x, y)
        for gname, grouping in zip(GROUPING_IDS, GROUPINGS):
            for sname, sub in (("full", None), ("sub10", [1, 0])):
                pick = (lambda t, axis: t) if sub is None else (lambda t, axis: t.index_select(axis, torch.tensor(sub)))
                case = {}

                def attach(ggn_field=None, grad=False, as_closures=False):
                    for p, V, g in zip(params, sqrt_ggn, grad_batch):
                        if ggn_field is not None:
                            V_sub = pick(V, 1).clone()
                            setattr(p, ggn_field, closures(V_sub) if as_closures else V_sub)
                        if grad:
                            p.grad_batch = pick(g, 0).clone()

                # -- EigvalshComputation (eigvalsh.py:79-225)
                groups = grouping(model)
                comp = vivit.EigvalshComputation(subsampling=sub)
                attach(comp._savefield, as_closures=True)
                run_hook(model, x, comp.get_extension_hook(groups))
                case["eigvalsh"] = [comp.get_result(g).clone() for g in groups]

                # -- EighComputation (eigh.py:101-275), criterion keep_nonzero (test/linalg/settings.py:23-44)
                groups = grouping(model, criterion=keep_nonzero)
                comp = vivit.EighComputation(subsampling=sub, warn_small_eigvals=0.0)
                attach(comp._savefield, as_closures=True)
                run_hook(model, x, comp.get_extension_hook(groups))
                res = [comp.get_result(g) for g in groups]
                case["eigh_evals"] = [r[0].clone() for r in res]
                case["eigh_evecs"] = [torch.cat([e.flatten(1) for e in r[1]], 1) for r in res]

                # -- DirectionalDerivativesComputation (directional_derivatives.py:216-353)
                groups = grouping(model, criterion=make_top_k(10), damping=constant_damping(1.0))
                comp = vivit.DirectionalDerivativesComputation(
                    subsampling_grad=sub, subsampling_ggn=sub, warn_small_eigvals=0.0
                )
                attach("sqrt_ggn_exact", grad=True)
                run_hook(model, x, comp.get_extension_hook(groups))
                res = [comp.get_result(g) for g in groups]
                case["gammas_abs"] = [r[0].abs().clone() for r in res]
                case["lambdas"] = [r[1].clone() for r in res]

                # -- DirectionalDampedNewtonComputation (directional_damped_newton.py:263-379)
                comp = vivit.DirectionalDampedNewtonComputation(
                    subsampling_grad=sub, subsampling_ggn=sub, warn_small_eigvals=0.0
                )
                attach("sqrt_ggn_exact", grad=True)
                run_hook(model, x, comp.get_extension_hook(groups))
                case["newton"] = [torch.cat([s.flatten() for s in comp.get_result(g)]) for g in groups]

                for p in params:  # the reference's hooks delete what they consumed
                    for field in ("sqrt_ggn_exact", "grad_batch", "vivit_ggn_exact"):
                        assert not hasattr(p, field), field
                out[(problem.name, gname, sname)] = case
        print(problem.name, "done", flush=True)
    out["__gram_hooks__"] = gram_hooks(vivit)
    out["__lanczos__"] = lanczos_vectors()
    out["__linear_closures__"] = linear_closures()
    out["__eig_utils__"] = eig_utils()
    out["__gram_utils__"] = gram_utils()
    out["__meta__"] = {"batch_sizes": {p.name: p.make()[2].shape[0] for p in PROBLEMS}, "torch": str(torch.__version__)}
    torch.save(out, os.path.join(HERE, "reference_run.pt"))
    print(f"wrote {len(out) - 6} cases + gram hooks + lanczos + linear closures + eig / gram / ggn utils")


if __name__ == "__main__":
    main()
