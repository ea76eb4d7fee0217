"""GGN eigenvalues and eigenvectors during back-propagation (``vivit/linalg/eigh.py``)."""

from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple
from warnings import warn

import torch
from torch import Tensor
from torch.nn import Module, Parameter

from vivit_b200 import kernels
from vivit_b200.factors import fold_linear_bias
from vivit_b200.linalg.eigvalsh import _accumulate_gram, _make_dist
from vivit_b200.linalg.solve_queue import SolveQueue
from vivit_b200.linalg.utils import get_hook_store_batch_size, get_vivit_extension, normalize
from vivit_b200.utils import delete_savefield, keep_indices
from vivit_b200.utils.checks import check_key_exists, check_subsampling_unique, check_unique_params
from vivit_b200.utils.hooks import ParameterGroupsHook


class EighComputation:
    """Provide the extension and hook to compute GGN eigenpairs (``vivit/linalg/eigh.py:21``).

    The loss must use ``reduction='mean'``.  ``process_group`` (not in the reference) shards
    the parameter dimension over the ranks of a ``torch.distributed`` group: eigenvectors are
    then returned as this rank's dim-0 slice of every parameter unless ``gather=True``.
    """

    def __init__(
        self,
        subsampling: Optional[List[int]] = None,
        mc_samples: int = 0,
        verbose: bool = False,
        warn_small_eigvals: float = 1e-4,
        process_group=None,
        gather: bool = False,
        batch_solves: bool = True,
        solve_queue: Optional[SolveQueue] = None,
    ):
        """``batch_solves`` (not in the reference): with several block-diagonal groups, the Gram matrices are
        decomposed in ONE batched solver call once the last group has fired (the groups are independent,
        ``vivit/utils/hooks.py:214-219``, and share ``R = C * N``); a group's factors stay alive until then.
        ``False`` restores the reference's order (solve and free inside every group's hook).

        ``solve_queue`` (not in the reference): a ``SolveQueue`` shared with other Computations.  The Gram
        matrices are handed to it and decomposed, together with everything else in the queue, when the first
        result is asked for (``linalg/solve_queue.py``)."""
        check_subsampling_unique(subsampling)
        self._subsampling = subsampling
        self._mc_samples = mc_samples
        self._verbose = verbose
        self._dist = _make_dist(process_group)
        self._gather = gather
        self._batch_solves = batch_solves
        self._shared_queue = solve_queue
        self._queue = solve_queue if solve_queue is not None else SolveQueue()
        self._savefield = self.get_extension().savefield
        self._warn_small_eigvals = warn_small_eigvals
        self._mc_state = None
        # filled during the backward pass, keys are group ids
        self._batch_size: Dict[int, int] = {}
        self._evals: Dict[int, Tensor] = {}
        self._evecs: Dict[int, List[Tensor]] = {}

    def get_result(self, group: Dict) -> Tuple[Tensor, List[Tensor]]:
        """``(evals [K], [evecs_p [K, *p.shape]])`` of a GGN block (``eigh.py:65-90``)."""
        gid = id(group)
        if gid not in self._evals and len(self._queue):
            self._queue.flush()
        try:
            return self._evals[gid], self._evecs[gid]
        except KeyError as e:
            raise KeyError("No results available for this group") from e

    def get_extension(self):
        """Extension for the ``with backpack(...)`` context (``eigh.py:92-99``)."""
        ext = get_vivit_extension(self._subsampling, self._mc_samples, self._dist.shard)
        ext.mc_state = getattr(self, "_mc_state", None)
        return ext

    def get_extensions(self):
        """Plural form named by the north star; a one-element list."""
        return [self.get_extension()]

    def get_extension_hook(self, param_groups: List[Dict]) -> Callable[[Module], None]:
        """Hook computing eigenpairs during back-propagation (``eigh.py:101-164``).

        Each group needs ``'params'`` and a ``'criterion'``: ``Callable[[Tensor], List[int]]``
        receiving the ascending eigenvalues and returning the indices to keep.
        """
        self._check_param_groups(param_groups)
        hook_store_batch_size = get_hook_store_batch_size(
            param_groups, self._batch_size, verbose=self._verbose
        )
        batch_sizes, subsampling, savefield = self._batch_size, self._subsampling, self._savefield
        evals, evecs, verbose = self._evals, self._evecs, self._verbose
        warn_small_eigvals, dist, gather = self._warn_small_eigvals, self._dist, self._gather
        queue, shared = self._queue, self._shared_queue is not None
        batch, fired = self._batch_solves and len(param_groups) > 1, []
        if not shared:
            queue.clear()  # leftovers of a backward pass that did not reach its last group

        def param_computation(hook: ParameterGroupsHook, param: Parameter) -> None:
            pass  # nothing per parameter (eigh.py:174-181)

        def accumulate(hook: ParameterGroupsHook, existing: None, update: None) -> None:
            pass  # (eigh.py:191-200)

        def group_hook(hook: ParameterGroupsHook, accumulation: None, group: Dict[str, Any]) -> None:
            """Gram matrix -> eigendecomposition -> filter -> parameter space (eigh.py:222-275)."""
            gid = id(group)
            if verbose:
                print(f"Group {gid}: Delete 'batch_size'")
            batch_size = batch_sizes.pop(gid)

            factors = [getattr(p, savefield)["_factor"] for p in group["params"]]
            members = {id(p) for p in group["params"]}
            for factor in factors:  # Linear weight + bias in one group: one structured Gram call (factors.py)
                fold_linear_bias(factor, lambda q: id(q) in members)
            gram = None
            for factor in factors:  # eigh.py:239-242
                gram = _accumulate_gram(gram, factor)
            C, N_ggn = factors[0].C, factors[0].N
            # eigh.py:245-246; over several ranks the rescale rides on the all-reduce of the partial Grams
            dist.scale_allreduce_(1.0 if subsampling is None else batch_size / len(subsampling), gram)

            queue.submit(gram, lambda gram_evals, gram_evecs: finish(group, gram_evals, gram_evecs, factors), dist=dist)
            fired.append(gid)
            if not shared and (not batch or len(fired) == len(param_groups)):
                del fired[:]
                queue.flush()  # eigh.py:248 (all groups of this pass in one call)

        def finish(group, gram_evals, gram_evecs, factors) -> None:
            gid = id(group)
            keep = group["criterion"](gram_evals)  # eigh.py:252-253
            keep_idx = keep_indices(keep, gram_evals)
            gram_evals = gram_evals.index_select(0, keep_idx)
            if warn_small_eigvals and (gram_evals.abs() < warn_small_eigvals).any():
                warn(
                    "Some eigenvectors have small eigenvalues."
                    + " Their parameter space transformation is numerically unstable."
                    + " This can spoil orthogonality of eigenvectors."
                    + " Maybe use a more restrictive eigenvalue filter criterion."
                )
            # eigenvectors selectable via first axis: [K, C*N]  (eigh.py:265)
            U = gram_evecs.index_select(1, keep_idx).t().contiguous()
            K = U.shape[0]

            norm2 = torch.zeros(K, dtype=torch.float64, device=U.device)
            group_evecs = []
            for param, factor in zip(group["params"], factors):  # eigh.py:267-270
                group_evecs.append(factor.backtransform(U, norm2))
                delete_savefield(param, savefield, verbose=verbose)
            dist.allreduce_(norm2)
            normalize(group_evecs, norm2)  # eigh.py:272
            if gather:
                group_evecs = [
                    dist.allgather_dim0(e, p.shape[0]) for e, p in zip(group_evecs, group["params"])
                ]
            evals[gid] = gram_evals
            evecs[gid] = group_evecs

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
        """Groups need ``'params'`` and ``'criterion'``; parameters are unique (``eigh.py:280-292``)."""
        check_key_exists(param_groups, "params")
        check_key_exists(param_groups, "criterion")
        check_unique_params(param_groups)
