"""GGN eigenvalues during back-propagation (``vivit/linalg/eigvalsh.py``)."""

from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
from torch import Tensor
from torch.nn import Module, Parameter

from vivit_b200 import kernels
from vivit_b200.factors import fold_linear_bias
from vivit_b200.linalg.solve_queue import SolveQueue
from vivit_b200.linalg.utils import get_hook_store_batch_size, get_vivit_extension
from vivit_b200.utils import delete_savefield
from vivit_b200.utils.checks import check_key_exists, check_subsampling_unique, check_unique_params
from vivit_b200.utils.hooks import ParameterGroupsHook


class EigvalshComputation:
    """Provide the extension and hook to compute GGN eigenvalues
    (``vivit/linalg/eigvalsh.py:20``).  The loss must use ``reduction='mean'``."""

    def __init__(
        self,
        subsampling: Optional[List[int]] = None,
        mc_samples: int = 0,
        verbose: bool = False,
        process_group=None,
        batch_solves: bool = True,
        solve_queue: Optional[SolveQueue] = None,
    ):
        """``batch_solves`` / ``solve_queue`` (not in the reference) as for ``EighComputation``: the Gram matrices of
        several block-diagonal groups are decomposed by one batched call once the last group has fired, or -- with
        a shared ``SolveQueue`` -- together with those of other Computations when the first result is read."""
        check_subsampling_unique(subsampling)
        self._subsampling = subsampling
        self._mc_samples = mc_samples
        self._verbose = verbose
        self._dist = _make_dist(process_group)
        self._batch_solves = batch_solves
        self._shared_queue = solve_queue
        self._queue = solve_queue if solve_queue is not None else SolveQueue()
        self._savefield = self.get_extension().savefield
        self._mc_state = None  # tests may pin the MC draw
        # filled during the backward pass, keys are group ids
        self._batch_size: Dict[int, int] = {}
        self._evals: Dict[int, Tensor] = {}

    def get_result(self, group: Dict) -> Tensor:
        """Ascending Gram eigenvalues of the group's GGN block (``eigvalsh.py:53-68``)."""
        if id(group) not in self._evals and len(self._queue):
            self._queue.flush()
        try:
            return self._evals[id(group)]
        except KeyError as e:
            raise KeyError("No results available for this group") from e

    def get_extension(self):
        """Extension for the ``with backpack(...)`` context (``eigvalsh.py:70-77``)."""
        ext = get_vivit_extension(self._subsampling, self._mc_samples, self._dist.shard)
        ext.mc_state = getattr(self, "_mc_state", None)
        return ext

    def get_extensions(self):
        """Plural form named by the north star; a one-element list."""
        return [self.get_extension()]

    def get_extension_hook(self, param_groups: List[Dict]) -> Callable[[Module], None]:
        """Hook computing the eigenvalues during back-propagation (``eigvalsh.py:79-131``)."""
        self._check_param_groups(param_groups)
        hook_store_batch_size = get_hook_store_batch_size(
            param_groups, self._batch_size, verbose=self._verbose
        )
        savefield, verbose = self._savefield, self._verbose
        subsampling, batch_sizes, evals, dist = self._subsampling, self._batch_size, self._evals, self._dist
        queue, shared = self._queue, self._shared_queue is not None
        batch, fired = self._batch_solves and len(param_groups) > 1, []
        if not shared:
            queue.clear()  # leftovers of a backward pass that did not reach its last group

        def param_computation(hook: ParameterGroupsHook, param: Parameter):
            # eager: evaluate this parameter's Gram and drop its factor (eigvalsh.py:145-158)
            factor = getattr(param, savefield)["_factor"]
            group = hook._group_of[id(param)]
            fold_linear_bias(factor, lambda q: hook._group_of.get(id(q)) is group)  # factors.py
            delete_savefield(param, savefield, verbose=verbose)
            return factor

        def accumulate(hook: ParameterGroupsHook, existing, update):
            # (eigvalsh.py:170-183): G_existing + G_update, done in place on one buffer
            if not isinstance(existing, Tensor):
                existing = _accumulate_gram(None, existing)
            return _accumulate_gram(existing, update)

        def group_hook(hook: ParameterGroupsHook, accumulation, group: Dict):
            gid = id(group)
            if verbose:
                print(f"Group {gid}: Delete 'batch_size'")
            batch_size = batch_sizes.pop(gid)
            gram = accumulation if isinstance(accumulation, Tensor) else _accumulate_gram(None, accumulation)
            # eigvalsh.py:218-219; over several ranks the rescale rides on the all-reduce of the partial Grams
            dist.scale_allreduce_(1.0 if subsampling is None else batch_size / len(subsampling), gram)

            def store(gram_evals, _):
                if verbose:
                    print(f"Group {gid}: Store 'gram_evals'")
                evals[gid] = gram_evals

            queue.submit(gram, store, vectors=False, dist=dist)  # eigvalsh.py:221
            fired.append(gid)
            if not shared and (not batch or len(fired) == len(param_groups)):
                del fired[:]
                queue.flush()

        hook = ParameterGroupsHook.from_functions(param_groups, param_computation, group_hook, accumulate)

        def extension_hook(module: Module) -> None:
            if verbose:
                print(f"Extension hook on module {id(module)} {module}")
            hook_store_batch_size(module)
            hook(module)

        if verbose:
            print("ID map groups → params")
            for group in param_groups:
                print(f"{id(group)} → {[id(p) for p in group['params']]}")
        return extension_hook

    @staticmethod
    def _check_param_groups(param_groups: List[Dict]) -> None:
        check_key_exists(param_groups, "params")
        check_unique_params(param_groups)


def _accumulate_gram(gram: Optional[Tensor], factor) -> Tensor:
    if gram is None:
        like = factor._like()
        gram = torch.zeros(factor.R, factor.R, dtype=like.dtype, device=like.device)
    factor.gram_accum(gram)
    return gram


def _make_dist(process_group):
    from vivit_b200.dist import ShardedReduce

    return ShardedReduce(process_group)
