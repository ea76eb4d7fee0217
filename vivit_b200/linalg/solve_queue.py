"""Deferred, batched eigendecompositions shared by several Computations (not in the reference).

The reference decomposes a group's Gram matrix inside the hook that assembled it (``vivit/linalg/eigh.py:248``,
``vivit/optim/directional_derivatives.py:291``).  The solver of this package is latency-bound at the sizes of the
path (``R = C * N`` of a thousand or so: a round of the block Jacobi is a chain of barriers and reductions that
keeps ~120 of the 148 SMs busy with very little work each), and a batch of such problems costs far less than the
sum of its members: two ``R = 1280`` matrices take 16.9 ms together and 13.2 ms each.  A ``SolveQueue`` collects
the matrices of every Computation it was handed to and decomposes them in ONE ``vvt_syevj_batched`` call when the
first result is asked for:

    queue = SolveQueue()
    eigh = EighComputation(solve_queue=queue)
    dirs = DirectionalDerivativesComputation(solve_queue=queue)
    with backpack(eigh.get_extension(), extension_hook=eigh.get_extension_hook(groups)):
        loss_fn(model(x), y).backward()
    with backpack(*dirs.get_extensions(), extension_hook=dirs.get_extension_hook(groups)):
        loss_fn(model(x), y).backward()
    evals, evecs = eigh.get_result(groups[0])      # <- both Gram matrices are decomposed here, together
    gammas, lambdas = dirs.get_result(groups[0])

Results are those of the immediate order (same kernels, and the batched solver treats its problems
independently); what changes is when they exist and that a group's factors stay alive until then.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor

from vivit_b200 import kernels


class SolveQueue:
    """Symmetric eigenproblems waiting for one batched solve."""

    def __init__(self) -> None:
        self._items: List[Tuple[Tensor, Callable[[Tensor, Optional[Tensor]], None], bool, object]] = []

    def __len__(self) -> int:
        return len(self._items)

    def clear(self) -> None:
        """Forget what was submitted (without solving)."""
        self._items = []

    def submit(self, gram: Tensor, done: Callable[[Tensor, Optional[Tensor]], None], vectors: bool = True,
               dist=None) -> None:
        """``done(evals, evecs)`` is called by ``flush`` with the ascending eigenvalues ``[R]`` and the
        eigenvectors ``[R, R]`` (columns; ``None`` with ``vectors=False``) of ``gram``.  ``gram`` must not be
        modified until then.  ``dist``: the ``ShardedReduce`` of a parameter-sharded Computation -- every rank
        submits the same (all-reduced) matrix, and a matrix that is solved on its own has the rounds of the
        two-level solver distributed over the ranks (``kernels.syevj_dist``)."""
        self._items.append((gram, done, vectors, dist))

    def flush(self) -> None:
        """Decompose everything submitted so far -- matrices of one shape, dtype and device (and one kind of request: with or without eigenvectors) in one batched
        call --
        and run the callbacks in submission order."""
        items, self._items = self._items, []
        if not items:
            return
        solved: List = [None] * len(items)
        buckets = {}
        for i, (gram, _, vectors, _) in enumerate(items):
            buckets.setdefault((tuple(gram.shape), gram.dtype, gram.device, vectors), []).append(i)
        for (_, _, _, vectors), members in buckets.items():
            if len(members) == 1:
                gram, dist = items[members[0]][0], items[members[0]][3]
                comm = dist.solver_comm(gram) if dist is not None else 0
                if comm:
                    solved[members[0]] = kernels.syevj_dist(
                        comm, dist.world, gram, vectors=vectors, p2p=dist.solver_arena(gram)
                    )
                else:
                    solved[members[0]] = kernels.syevj(gram, vectors=vectors)
                continue
            evals, evecs = kernels.syevj_batched(torch.stack([items[i][0] for i in members]), vectors=vectors)
            for slot, i in enumerate(members):
                solved[i] = (evals[slot], evecs[slot] if vectors else None)
        # every callback runs, also after one of them has raised (a criterion of one Computation must not cost the
        # others their results); the first exception is re-raised at the end
        failed = None
        for (_, done, _, _), (evals, evecs) in zip(items, solved):
            try:
                done(evals, evecs)
            except Exception as e:  # noqa: BLE001
                failed = failed or e
        if failed is not None:
            raise failed
