"""Directionally damped Newton steps (``vivit/optim/directional_damped_newton.py``)."""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from vivit_b200 import kernels
from vivit_b200.backprop.extensions import BatchGrad
from vivit_b200.linalg.eigvalsh import _make_dist
from vivit_b200.linalg.solve_queue import SolveQueue
from vivit_b200.linalg.utils import get_hook_store_batch_size
from vivit_b200.optim.directional_derivatives import DirectionalDerivativesComputation as _DD
from vivit_b200.optim.utils import get_sqrt_ggn_extension
from vivit_b200.utils.checks import check_key_exists, check_subsampling_unique, check_unique_params
from vivit_b200.utils.hooks import ParameterGroupsHook


class DirectionalDampedNewtonComputation:
    r"""Provide extensions and hook for directionally damped Newton steps
    :math:`s = \sum_k -\gamma_k / (\lambda_k + \delta_k) e_k`
    (``vivit/optim/directional_damped_newton.py:24``).  The loss must use ``reduction='mean'``.
    """

    def __init__(
        self,
        subsampling_grad: Optional[List[int]] = None,
        subsampling_ggn: Optional[List[int]] = None,
        mc_samples_ggn: Optional[int] = 0,
        verbose: Optional[bool] = False,
        warn_small_eigvals: float = 1e-4,
        process_group=None,
        gather: bool = False,
        solve_queue: Optional[SolveQueue] = None,
    ):
        """``solve_queue`` (not in the reference): a ``SolveQueue`` shared with other Computations; the group's Gram
        matrix is decomposed, together with everything else in the queue, when the first result is asked for
        (``linalg/solve_queue.py``).  The group's factors stay alive until then (the step applies ``V``)."""
        check_subsampling_unique(subsampling_grad)
        self._queue = solve_queue
        check_subsampling_unique(subsampling_ggn)
        self._mc_samples_ggn = mc_samples_ggn
        if self._mc_samples_ggn != 0:
            assert mc_samples_ggn == 1  # directional_damped_newton.py:81-82
        self._subsampling_grad = subsampling_grad
        self._subsampling_ggn = subsampling_ggn
        self._dist = _make_dist(process_group)
        self._gather = gather
        self._savefield_grad = BatchGrad.savefield
        self._savefield_ggn = get_sqrt_ggn_extension(None, mc_samples_ggn).savefield
        self._verbose = verbose
        self._warn_small_eigvals = warn_small_eigvals
        self._mc_state = None
        # filled during the backward pass, keys are group ids
        self._batch_size: Dict[int, int] = {}
        self._newton_steps: Dict[int, Tuple[Tensor]] = {}

    def get_result(self, group: Dict) -> Tuple[Tensor]:
        """Damped Newton step in the format of ``group['params']``
        (``directional_damped_newton.py:101-120``)."""
        if id(group) not in self._newton_steps and self._queue is not None:
            self._queue.flush()
        try:
            return self._newton_steps[id(group)]
        except KeyError as e:
            raise KeyError("No results available for this group") from e

    def get_extensions(self) -> List:
        """``[BatchGrad, SqrtGGN{Exact,MC}]`` (``directional_damped_newton.py:122-135``)."""
        grad = BatchGrad(subsampling=self._subsampling_grad, lazy=True)
        ggn = get_sqrt_ggn_extension(self._subsampling_ggn, self._mc_samples_ggn, lazy=True)
        grad._shard = ggn._shard = self._dist.shard
        ggn.mc_state = getattr(self, "_mc_state", None)
        return [grad, ggn]

    def get_extension_hook(self, param_groups: List[Dict]) -> Callable[[Module], None]:
        """Hook computing the Newton step during back-propagation
        (``directional_damped_newton.py:137-223``).  Groups need ``'params'``, ``'criterion'`` and
        ``'damping'``: ``Callable[[evals, gram_evecs, gammas, lambdas], Tensor[K]]``."""
        self._check_param_groups(param_groups)
        hook_store_batch_size = get_hook_store_batch_size(
            param_groups, self._batch_size, verbose=self._verbose
        )
        factors: Dict[int, object] = {}  # the factors outlive param_computation (:258 keeps V)

        def param_computation(hook, param):
            V, g = _DD._param_computation(
                hook, param, self._savefield_ggn, self._savefield_grad, self._verbose, True
            )
            factors[id(param)] = V
            return (V, g)

        hook = ParameterGroupsHook.from_functions(
            param_groups,
            param_computation,
            lambda hook, accumulation, group: self._group_hook(
                hook, accumulation, group, self._batch_size, factors, self._newton_steps,
                self._verbose, self._warn_small_eigvals, self._dist, self._gather, self._queue,
            ),
            lambda hook, existing, update: _DD._accumulate(hook, existing, update, self._verbose),
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
    def _group_hook(hook, accumulation, group, batch_size, factors, newton_steps, verbose,
                    warn_small_eigvals, dist, gather, queue=None):
        """Directions, directional derivatives, dampings, coefficients in Gram space, then one
        application of ``V`` per parameter (``directional_damped_newton.py:263-379``)."""
        gid = id(group)
        N = batch_size.pop(gid)
        acc, gram, corr2 = _DD._gram_space(accumulation, N, dist)

        def finish(evals, evecs):
            evals, evecs, gammas, lambdas = _DD._filter_and_evaluate(
                acc, evals, evecs, group, N, verbose, warn_small_eigvals
            )
            deltas = group["damping"](evals, evecs, gammas, lambdas)  # :353
            deltas = torch.as_tensor(deltas, dtype=evals.dtype, device=evals.device).reshape(-1)
            # coefficients, weighting in Gram space and the V_correction rescale (:354-366)
            v = kernels.newton_coeff(evecs, gammas, lambdas, deltas, evals, math.sqrt(corr2))
            steps = []
            for param in group["params"]:  # :368-377
                step = factors.pop(id(param)).v_apply(v)
                if gather:
                    step = dist.allgather_dim0(step[None], param.shape[0])[0]
                steps.append(step)
            newton_steps[gid] = steps

        if queue is not None:
            queue.submit(gram, finish, dist=dist)
            return
        if verbose:
            print(f"Group {gid}: Eigen-decompose Gram matrix")
        finish(*kernels.syevj(gram, vectors=True))  # :315

    @staticmethod
    def _check_param_groups(param_groups: List[Dict]) -> None:
        """(``directional_damped_newton.py:410-419``)."""
        check_key_exists(param_groups, "params")
        check_key_exists(param_groups, "criterion")
        check_key_exists(param_groups, "damping")
        check_unique_params(param_groups)
