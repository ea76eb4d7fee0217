"""Parameter-sharded Gram assembly across the GPUs of one box.

``G = sum_p V_p^T V_p`` and ``V^T g = sum_p V_p^T g_p`` are sums over parameter entries,
so each rank takes a slice of every parameter's leading (output-channel) dimension,
assembles a partial Gram from its slice, and the partial Grams are summed with ONE
all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  Everything after the
all-reduce that lives in Gram space (directional derivatives, Newton coefficients, the
eigensolver of small groups) is computed redundantly and deterministically on every rank; the
eigensolver of a group of 4096 columns or more (fp32) is distributed over the ranks
(``solver_comm`` / ``solver_arena``, ``vvt_syevj_dist``); results in
parameter space (eigenvectors, Newton steps) stay sharded along dim 0 unless
``gather=True``.

The factor back-propagation itself is replicated (activations and ``S`` are
``O(R x width)``).  Block-diagonal groups need no communication at all; they can be
assigned to ranks whole (``assign_groups``).
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import ctypes

import torch
from torch import Tensor


_ARENA_OWNER: Dict[int, object] = {}  # device index -> process group its peer-memory arena is mapped for
_NO_ARENA = object()


def shard_bounds(size: int, rank: int, world: int) -> Tuple[int, int]:
    """``[lo, hi)`` of ``range(size)`` owned by ``rank`` (balanced, contiguous)."""
    return (size * rank) // world, (size * (rank + 1)) // world


class ShardedReduce:
    """All-reduce plumbing for one process group (``None`` = single process, no-ops)."""

    def __init__(self, process_group=None, gather: bool = False):
        self.group = process_group
        self.gather = gather
        if process_group is None:
            self.rank, self.world = 0, 1
        else:
            import torch.distributed as dist

            self.rank = dist.get_rank(process_group)
            self.world = dist.get_world_size(process_group)

    @property
    def shard(self) -> Optional[Tuple[int, int]]:
        return None if self.world == 1 else (self.rank, self.world)

    def _nccl_comm(self, like: Tensor) -> int:
        """Raw ``ncclComm_t`` of the group for ``vvt_nccl_allreduce_gram`` (0: not an NCCL group, or this torch
        build does not hand the communicator out -- then ``torch.distributed`` carries the collective)."""
        if not like.is_cuda:
            return 0
        if getattr(self, "_comm", None) is None:
            ptr = 0
            try:
                backend = self.group._get_backend(torch.device("cuda"))
                ptr = int(backend._comm_ptr())
            except Exception:
                ptr = 0
            self._comm = ptr
        return self._comm

    def solver_comm(self, gram: Tensor) -> int:
        """``ncclComm_t`` for the distributed eigensolver rounds (``vvt_syevj_dist``), 0 = solve on every rank:
        one process, not an NCCL group, or a problem off the two-level path (fp64, fewer than 4096 columns), which
        every rank solves for itself anyway.  ``VVT_SYEVJ_DIST=0`` switches the distribution off."""
        import os

        if self.world == 1 or gram.dtype != torch.float32 or gram.shape[0] < int(os.environ.get("VVT_SYEVJ_WIDE_MIN", 4096)):
            return 0
        if os.environ.get("VVT_SYEVJ_DIST", "1") == "0":
            return 0
        return self._nccl_comm(gram)

    def solver_arena(self, gram: Tensor) -> bool:
        """Peer-memory arena for the distributed solver (``vvt_dist_arena_*``): every rank allocates one, the CUDA IPC
        handles are all-gathered over this group and the peers' arenas mapped, after which blocks change owner by
        direct stores over NVLink instead of ``ncclSend`` / ``ncclRecv``.  Collective; ``True`` when the arena of this
        process is mapped for THIS group and holds an ``R``-column factor.  A process has one arena per device: a
        second group on the same device keeps the NCCL hand-over.  ``VVT_SYEVJ_P2P=0`` switches it off."""
        import os
        import warnings

        import torch.distributed as dist

        from vivit_b200 import _lib

        if os.environ.get("VVT_SYEVJ_P2P", "1") == "0" or self.world == 1 or self.world > 16 or not gram.is_cuda:
            return False
        dev = gram.device.index if gram.device.index is not None else torch.cuda.current_device()
        bound = _ARENA_OWNER.get(dev)
        if bound is not None and bound is not self.group:
            return False  # also after a failed attempt (bound to the sentinel)
        lib = _lib.load()
        need = int(lib.vvt_dist_arena_bytes_for(int(gram.shape[0])))
        with torch.cuda.device(gram.device):
            if bound is self.group and int(lib.vvt_dist_arena_bytes()) >= need:
                return True
            ok = True
            try:
                if bound is self.group:  # grow: nobody may write into the old arena any more
                    torch.cuda.synchronize()
                    _lib.check(lib.vvt_dist_arena_close_peers(), "vvt_dist_arena_close_peers")
                    dist.barrier(group=self.group)
                    _lib.check(lib.vvt_dist_arena_free(), "vvt_dist_arena_free")
                handle = (ctypes.c_ubyte * 64)()
                _lib.check(lib.vvt_dist_arena_alloc(need, handle), "vvt_dist_arena_alloc")
            except Exception as e:  # noqa: BLE001
                ok, why = False, str(e)
            mine = torch.tensor(list(bytes(handle)) if ok else [0] * 64, dtype=torch.uint8, device=gram.device)
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=gram.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()):
                every = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(every, mine, group=self.group)
                handles = b"".join(t.cpu().numpy().tobytes() for t in every)
                try:
                    _lib.check(lib.vvt_dist_arena_open(handles, self.world, self.rank), "vvt_dist_arena_open")
                except Exception as e:  # noqa: BLE001
                    ok, why = False, str(e)
                flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=gram.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()):
                _ARENA_OWNER[dev] = self.group
                return True
            # some rank could not export or map its arena: every rank drops its own and keeps the NCCL hand-over
            lib.vvt_dist_arena_close_peers()
            dist.barrier(group=self.group)
            lib.vvt_dist_arena_free()
            _ARENA_OWNER[dev] = _NO_ARENA
            warnings.warn(
                "vivit_b200: CUDA IPC peer mapping is not available ("
                + (why if not ok else "on another rank")
                + "); the distributed eigensolver hands blocks over with ncclSend / ncclRecv"
            )
            return False

    def scale_allreduce_(self, alpha: float, gram: Tensor, cross: Optional[Tensor] = None) -> None:
        """``gram`` (and ``cross``) ``<- alpha * sum over ranks`` in place: the partial-Gram exchange of the
        parameter-sharded path with the sub-sampling rescale of ``eigh.py:245-246`` folded in.  One process:
        just the rescale."""
        from vivit_b200 import kernels

        if self.world == 1:
            if alpha != 1.0:
                kernels.scale_(gram, alpha)
                if cross is not None:
                    kernels.scale_(cross, alpha)
            return
        comm = self._nccl_comm(gram)
        if comm:
            kernels.nccl_allreduce_gram(comm, gram, cross, alpha)  # one NCCL group, pre-multiplied sum
            return
        if alpha != 1.0:
            kernels.scale_(gram, alpha)
            if cross is not None:
                kernels.scale_(cross, alpha)
        self.allreduce_(*([gram] if cross is None else [gram, cross]))

    def allreduce_(self, *tensors: Tensor) -> None:
        """Sum the given tensors over ranks, in place, with a single collective."""
        if self.world == 1 or not tensors:
            return
        import torch.distributed as dist

        if len(tensors) == 1:
            dist.all_reduce(tensors[0], group=self.group)
            return
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.all_reduce(flat, group=self.group)
        offset = 0
        for t in tensors:
            t.copy_(flat[offset : offset + t.numel()].reshape(t.shape))
            offset += t.numel()

    def allgather_dim0(self, t: Tensor, full_size: int) -> Tensor:
        """Concatenate dim-0 shards of a parameter-shaped tensor (leading stack axis kept)."""
        if self.world == 1:
            return t
        import torch.distributed as dist

        # shards may differ by one row: pad to the largest
        sizes = [shard_bounds(full_size, r, self.world) for r in range(self.world)]
        width = max(hi - lo for lo, hi in sizes)
        pad_shape = list(t.shape)
        pad_shape[1] = width
        buf = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        buf[:, : t.shape[1]] = t
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf, group=self.group)
        return torch.cat([o[:, : hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=1)


def assign_groups(costs: Sequence[float], world: int) -> List[int]:
    """Greedy longest-processing-time assignment of independent (block-diagonal) groups
    to ranks; returns the owning rank of each group."""
    loads = [0.0] * world
    owner = [0] * len(costs)
    for i in sorted(range(len(costs)), key=lambda i: -costs[i]):
        r = min(range(world), key=lambda r: loads[r])
        owner[i] = r
        loads[r] += costs[i]
    return owner


def team_sizes(costs: Sequence[float], world: int) -> List[int]:
    """Ranks per group when there are more ranks than groups: one each, the rest go one by one to the group whose
    cost per rank is largest."""
    sizes = [1] * len(costs)
    for _ in range(world - len(costs)):
        i = max(range(len(costs)), key=lambda i: costs[i] / sizes[i])
        sizes[i] += 1
    return sizes


def team_groups(param_groups: Sequence[Dict], process_group=None, costs: Optional[Sequence[float]] = None):
    """``local_groups`` for MORE ranks than block-diagonal groups: every group gets a team of consecutive ranks.
    A team of one runs its group as ``local_groups`` would; a larger team runs the Computation with
    ``process_group=team`` -- the group's Gram is parameter-sharded over the team (one all-reduce inside the team)
    and a Gram of 4096 columns or more is decomposed by the team together (``vvt_syevj_dist``).  Teams never talk to
    each other.

    Returns ``(own, team, teams)``: the groups this rank works on, the ``torch.distributed`` group of its team
    (``None`` for a team of one, or when there are at most as many ranks as groups -- then this is ``local_groups``),
    and the list of rank lists per group.  Collective over ``process_group`` (sub-groups are created)."""
    if process_group is None:
        return list(param_groups), None, [[0] for _ in param_groups]
    import torch.distributed as dist

    rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
    if world <= len(param_groups):
        own, owner = local_groups(param_groups, process_group, costs)
        return own, None, [[o] for o in owner]
    if costs is None:
        numel = [sum(p.numel() for p in g["params"]) for g in param_groups]
        total = float(max(sum(numel), 1))
        costs = [1.0 + n / total for n in numel]
    sizes = team_sizes(costs, world)
    teams, start = [], 0
    for size in sizes:
        teams.append(list(range(start, start + size)))
        start += size
    own, mine = [], None
    for group, members in zip(param_groups, teams):
        # every rank of the parent group takes part in the creation of every sub-group, in the same order
        global_ranks = [dist.get_global_rank(process_group, r) for r in members]
        sub = dist.new_group(global_ranks) if len(members) > 1 else None
        if rank in members:
            own.append(group)
            mine = sub
    return own, mine, teams


def local_groups(param_groups: Sequence[Dict], process_group=None, costs: Optional[Sequence[float]] = None):
    """Block-diagonal parameter groups are independent (``vivit/utils/hooks.py:214-219``): give every rank
    whole groups and run the Computations WITHOUT ``process_group`` on them -- no collective at all.

    Returns ``(own, owner)``: the sub-list of ``param_groups`` (the same dict objects, for ``get_result``)
    this rank is responsible for, and the owning rank of every group.  ``costs`` default to one unit per
    group (the eigensolver, which depends on ``C N`` only) plus the group's share of the parameters.
    """
    if process_group is None:
        return list(param_groups), [0] * len(param_groups)
    import torch.distributed as dist

    rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
    if costs is None:
        numel = [sum(p.numel() for p in g["params"]) for g in param_groups]
        total = float(max(sum(numel), 1))
        costs = [1.0 + n / total for n in numel]
    owner = assign_groups(costs, world)
    return [g for g, o in zip(param_groups, owner) if o == rank], owner
