"""1st- and 2nd-order directional derivatives along GGN eigenvectors
(``vivit/optim/directional_derivatives.py``)."""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple
from warnings import warn

import torch
from torch import Tensor
from torch.nn import Module

from vivit_b200 import kernels
from vivit_b200.backprop.extensions import BatchGrad
from vivit_b200.factors import Factor, GradFactor, fold_linear_bias
from vivit_b200.linalg.eigvalsh import _make_dist
from vivit_b200.linalg.solve_queue import SolveQueue
from vivit_b200.linalg.utils import get_hook_store_batch_size
from vivit_b200.optim.utils import get_sqrt_ggn_extension
from vivit_b200.utils import delete_savefield, keep_indices
from vivit_b200.utils.checks import check_key_exists, check_subsampling_unique, check_unique_params
from vivit_b200.utils.hooks import ParameterGroupsHook


class _GramSpace:
    """Accumulated Gram-space quantities of a group: ``V^T V`` and ``V^T g`` (un-rescaled)."""

    def __init__(self, factor: Factor, grad: GradFactor, n_grad: int):
        like = factor._like()
        self.C, self.N_ggn, self.n_grad = factor.C, factor.N, n_grad
        self.V_t_V = torch.zeros(factor.R, factor.R, dtype=like.dtype, device=like.device)
        self.V_t_g_n = torch.zeros(factor.R, n_grad, dtype=like.dtype, device=like.device)


class DirectionalDerivativesComputation:
    """Provide extensions and hook for directional derivatives along GGN eigenvectors
    (``vivit/optim/directional_derivatives.py:24``).  The loss must use ``reduction='mean'``.

    Unlike the reference, which materialises ``V`` and the per-sample gradients for every
    parameter (``:238-239``), Linear layers stay structured here and only their ``S``, ``Z``
    and output gradients are touched.
    """

    def __init__(
        self,
        subsampling_grad: Optional[List[int]] = None,
        subsampling_ggn: Optional[List[int]] = None,
        mc_samples_ggn: Optional[int] = 0,
        verbose: Optional[bool] = False,
        warn_small_eigvals: float = 1e-4,
        process_group=None,
        solve_queue: Optional[SolveQueue] = None,
    ):
        """``solve_queue`` (not in the reference): a ``SolveQueue`` shared with other Computations; the group's
        Gram matrix is decomposed, together with everything else in the queue, when the first result is asked
        for (``linalg/solve_queue.py``).  Only ``V^T V`` and ``V^T g`` wait for it, the factors are freed as
        in the immediate order."""
        check_subsampling_unique(subsampling_grad)
        check_subsampling_unique(subsampling_ggn)
        self._queue = solve_queue
        self._mc_samples_ggn = mc_samples_ggn
        if self._mc_samples_ggn != 0:
            assert mc_samples_ggn == 1  # directional_derivatives.py:73-74
        self._subsampling_grad = subsampling_grad
        self._subsampling_ggn = subsampling_ggn
        self._dist = _make_dist(process_group)
        self._savefield_grad = BatchGrad.savefield
        self._savefield_ggn = get_sqrt_ggn_extension(None, mc_samples_ggn).savefield
        self._verbose = verbose
        self._warn_small_eigvals = warn_small_eigvals
        self._mc_state = None
        # filled during the backward pass, keys are group ids
        self._batch_size: Dict[int, int] = {}
        self._gammas: Dict[int, Tensor] = {}
        self._lambdas: Dict[int, Tensor] = {}

    def get_result(self, group: Dict) -> Tuple[Tensor, Tensor]:
        """``(gammas [N_grad, K], lambdas [N_ggn, K])`` (``directional_derivatives.py:94-117``)."""
        gid = id(group)
        if gid not in self._gammas and self._queue is not None:
            self._queue.flush()
        try:
            return self._gammas[gid], self._lambdas[gid]
        except KeyError as e:
            raise KeyError("No results available for this group") from e

    def get_extensions(self) -> List:
        """``[BatchGrad, SqrtGGN{Exact,MC}]`` for the ``with backpack(...)`` context
        (``directional_derivatives.py:119-132``)."""
        grad = BatchGrad(subsampling=self._subsampling_grad, lazy=True)
        ggn = get_sqrt_ggn_extension(self._subsampling_ggn, self._mc_samples_ggn, lazy=True)
        grad._shard = ggn._shard = self._dist.shard
        ggn.mc_state = getattr(self, "_mc_state", None)
        return [grad, ggn]

    def get_extension_hook(self, param_groups: List[Dict]) -> Callable[[Module], None]:
        """Hook computing ``gamma`` and ``lambda`` during back-propagation
        (``directional_derivatives.py:134-214``)."""
        self._check_param_groups(param_groups)
        hook_store_batch_size = get_hook_store_batch_size(
            param_groups, self._batch_size, verbose=self._verbose
        )
        hook = ParameterGroupsHook.from_functions(
            param_groups,
            lambda hook, param: self._param_computation(
                hook, param, self._savefield_ggn, self._savefield_grad, self._verbose, True
            ),
            lambda hook, accumulation, group: self._group_hook(
                hook, accumulation, group, self._batch_size, self._gammas, self._lambdas,
                self._verbose, self._warn_small_eigvals, self._dist, self._queue,
            ),
            lambda hook, existing, update: self._accumulate(hook, existing, update, self._verbose),
        )

        def extension_hook(module: Module) -> None:
            if self._verbose:
                print(f"Extension hook on module {id(module)} {module}")
            hook_store_batch_size(module)
            hook(module)

        if self._verbose:
            print("ID map groups → params")
            for group in param_groups:
                print(f"{id(group)} → {[id(p) for p in group['params']]}")
        return extension_hook

    @staticmethod
    def _param_computation(hook, param, savefield_ggn, savefield_grad, verbose, free_factor):
        """Hand this parameter's factor and per-sample gradients to the accumulator
        (``directional_derivatives.py:216-252``)."""
        V = getattr(param, savefield_ggn)
        g = getattr(param, savefield_grad)
        group = hook._group_of[id(param)]
        fold_linear_bias(V, lambda q: hook._group_of.get(id(q)) is group)  # one Gram call per Linear layer
        if verbose:
            print(f"Param {id(param)}: Compute V_t_V and V_t_g_n")
        if free_factor:
            delete_savefield(param, savefield_ggn, verbose=verbose)
        delete_savefield(param, savefield_grad, verbose=verbose)
        return (V, g)

    @staticmethod
    def _accumulate(hook, existing, update, verbose):
        """In-place accumulation of the dot products (``directional_derivatives.py:328-353``)."""
        if not isinstance(existing, _GramSpace):
            existing = DirectionalDerivativesComputation._start(existing)
        V, g = update
        if verbose:
            print("Accumulate dot product V_t_V, V_t_g_n")
        V.gram_accum(existing.V_t_V)  # partial_contract(V, V, (2, 2))  :245
        V.cross_accum(existing.V_t_g_n, g)  # partial_contract(V, g, (2, 1))  :246
        return existing

    @staticmethod
    def _start(first) -> _GramSpace:
        V, g = first
        n_grad = g.Dl.shape[0] if hasattr(g, "Dl") else g.g.shape[0]
        acc = _GramSpace(V, g, n_grad)
        V.gram_accum(acc.V_t_V)
        V.cross_accum(acc.V_t_g_n, g)
        return acc

    @staticmethod
    def _gram_space(accumulation, N, dist):
        """Finished Gram-space quantities of a group and the matrix to decompose, ``corr^2 V^T V``
        (``directional_derivatives.py:281-291``)."""
        if not isinstance(accumulation, _GramSpace):
            accumulation = DirectionalDerivativesComputation._start(accumulation)
        acc = accumulation
        dist.scale_allreduce_(1.0, acc.V_t_V, acc.V_t_g_n)  # the one exchange of the sharded path
        corr2 = N / acc.N_ggn  # V_correction**2  (:285-287)
        # eigenpairs of corr^2 V^T V; the solver leaves its input intact, so scale a copy
        gram = acc.V_t_V if corr2 == 1.0 else kernels.scale_(acc.V_t_V.clone(), corr2)
        return acc, gram, corr2

    @staticmethod
    def _filter_and_evaluate(acc, evals, evecs, group, N, verbose, warn_small_eigvals):
        """Filter the directions, evaluate ``gamma`` and ``lambda`` (``directional_derivatives.py:293-325``)."""
        gid = id(group)
        keep = group["criterion"](evals)  # :293
        keep_idx = keep_indices(keep, evals)
        if verbose:
            print(f"Group {gid}: Filter directions ({len(evals)} → {keep_idx.numel()})")
        evals = evals.index_select(0, keep_idx)
        evecs = evecs.index_select(1, keep_idx).contiguous()  # [R, K]

        if warn_small_eigvals and (evals.abs() < warn_small_eigvals).any():
            warn(
                "Some eigenvalues are small. This can lead to numerical instabilities"
                + " in the directional gradients because they require division by the"
                + " eigenvalue square root."
                + " Maybe use a more restrictive eigenvalue filter criterion."
            )
        if verbose:
            print(f"Group {gid}: Compute gammas and lambdas")
        gammas, lambdas = kernels.dirderiv_epilogue(
            acc.V_t_V, acc.V_t_g_n, evecs, evals, acc.C, acc.N_ggn, N
        )  # :302-325
        return evals, evecs, gammas, lambdas

    @staticmethod
    def _directions(accumulation, group, N, verbose, warn_small_eigvals, dist):
        """Eigendecompose the Gram matrix, filter, evaluate ``gamma`` and ``lambda``
        (``directional_derivatives.py:281-325`` == ``directional_damped_newton.py:304-351``)."""
        cls = DirectionalDerivativesComputation
        acc, gram, corr2 = cls._gram_space(accumulation, N, dist)
        if verbose:
            print(f"Group {id(group)}: Eigen-decompose Gram matrix")
        evals, evecs = kernels.syevj(gram, vectors=True)  # :291
        evals, evecs, gammas, lambdas = cls._filter_and_evaluate(
            acc, evals, evecs, group, N, verbose, warn_small_eigvals
        )
        return evals, evecs, gammas, lambdas, math.sqrt(corr2), acc.C, acc.N_ggn

    @staticmethod
    def _group_hook(
        hook, accumulation, group, batch_size, gammas, lambdas, verbose, warn_small_eigvals, dist, queue=None
    ):
        """Store ``gamma[n, k]`` and ``lambda[n, k]`` of the group (``directional_derivatives.py:255-325``)."""
        cls = DirectionalDerivativesComputation
        gid = id(group)
        N = batch_size.pop(gid)
        if queue is None:
            _, _, gammas[gid], lambdas[gid], _, _, _ = cls._directions(
                accumulation, group, N, verbose, warn_small_eigvals, dist
            )
            return
        acc, gram, _ = cls._gram_space(accumulation, N, dist)

        def done(evals, evecs):
            _, _, gammas[gid], lambdas[gid] = cls._filter_and_evaluate(
                acc, evals, evecs, group, N, verbose, warn_small_eigvals
            )

        queue.submit(gram, done, dist=dist)

    @staticmethod
    def _check_param_groups(param_groups: List[Dict]) -> None:
        check_key_exists(param_groups, "params")
        check_key_exists(param_groups, "criterion")
        check_unique_params(param_groups)
