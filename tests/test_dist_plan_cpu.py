"""Host logic of the distributed eigensolver (``vvt_syevj_dist``) without a GPU: the block hand-over plan
of the round-robin tournament over several ranks (``vvt_dbg_dist_plan`` runs the planner the solver uses), and the
routing of a parameter-sharded Computation's Gram matrix to the distributed solve in ``SolveQueue.flush``."""
import ctypes

import pytest
import torch

from vivit_b200 import _lib, kernels
from vivit_b200.linalg.solve_queue import SolveQueue


def pairing(nbw, rnd):
    """Blocks of every pair of a round, restated from the solver (``rr_pair`` / ``wide_blocks``): the intra round
    pairs (2i, 2i + 1), round r of the tournament pairs (r, nbw - 1) and ((r + i) mod m, (r - i) mod m), m = nbw - 1."""
    if rnd < 0:
        return [(2 * i, 2 * i + 1) for i in range(nbw // 2)]
    m = nbw - 1
    return [(rnd, m)] + [((rnd + i) % m, (rnd - i) % m) for i in range(1, nbw // 2)]


def plan(nbw, world, rnd, owner):
    lib = _lib.load()
    own = (ctypes.c_int * nbw)(*owner)
    moves = (ctypes.c_int * (3 * nbw))()
    n = ctypes.c_int(0)
    ranks = (ctypes.c_int * (nbw // 2))()
    _lib.check(lib.vvt_dbg_dist_plan(nbw, world, rnd, own, moves, nbw, ctypes.byref(n), ranks), "vvt_dbg_dist_plan")
    triples = [tuple(moves[3 * i : 3 * i + 3]) for i in range(n.value)]
    return list(own), triples, list(ranks)


@pytest.mark.parametrize("nbw,world", [(8, 2), (18, 2), (18, 4), (64, 3), (160, 2), (160, 4), (160, 8), (6, 8)])
def test_block_hand_over_plan(nbw, world):
    pairs = nbw // 2
    owner = [-1] * nbw
    for sweep in range(2):
        for rnd in range(-1, nbw - 1):
            before = list(owner)
            owner, moves, pair_rank = plan(nbw, world, rnd, owner)
            blocks = pairing(nbw, rnd)
            assert sorted(b for ab in blocks for b in ab) == list(range(nbw))  # every block in exactly one pair
            # contiguous, balanced slices of the pair index
            assert pair_rank == sorted(pair_rank) and set(pair_rank) <= set(range(world))
            sizes = [pair_rank.count(r) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
            # ... and they are the slices the solver launches its kernels on: [P r / W, P (r + 1) / W)
            for r in range(world):
                lo, hi = pairs * r // world, pairs * (r + 1) // world
                assert [i for i, pr in enumerate(pair_rank) if pr == r] == list(range(lo, hi))
            # during the round both blocks of a pair are with the rank that works on the pair
            for (a, b), r in zip(blocks, pair_rank):
                assert owner[a] == r and owner[b] == r
            # a move takes a block from the rank that held it to the rank that needs it, once
            assert len({m[0] for m in moves}) == len(moves)
            for blk, src, dst in moves:
                assert before[blk] == src and owner[blk] == dst and src != dst
            moved = {m[0] for m in moves}
            for blk in range(nbw):
                if blk not in moved:
                    assert before[blk] in (-1, owner[blk])
            if sweep == 0 and rnd == -1:
                assert moves == []  # the initial factor is replicated
            if rnd >= 1 and pairs >= 2 * world:
                # between two rounds of the tournament a rank hands one block to each neighbour
                for r in range(world):
                    out = [m for m in moves if m[1] == r]
                    assert len(out) <= 2 and all(abs(m[2] - r) == 1 for m in out)
                    assert len([m for m in moves if m[2] == r]) <= 2


def test_one_rank_never_moves_a_block():
    owner = [-1] * 16
    for rnd in range(-1, 15):
        owner, moves, pair_rank = plan(16, 1, rnd, owner)
        assert moves == [] and set(pair_rank) == {0}


class _FakeDist:
    """Stands in for ``ShardedReduce`` on an NCCL group (no GPU here)."""

    world = 2

    def __init__(self, comm):
        self.comm, self.arena_asked = comm, 0

    def solver_comm(self, gram):
        return self.comm

    def solver_arena(self, gram):
        self.arena_asked += 1
        return True


def test_solve_queue_routes_sharded_matrices_to_the_distributed_solver(monkeypatch):
    calls = []

    def fake_dist(comm, world, G, vectors=True, return_info=False, p2p=False):
        calls.append(("dist", comm, world, tuple(G.shape), vectors, p2p))
        return torch.zeros(G.shape[0]), (torch.eye(G.shape[0]) if vectors else None)

    def fake_one(G, vectors=True, return_info=False):
        calls.append(("one", tuple(G.shape), vectors))
        return torch.zeros(G.shape[0]), (torch.eye(G.shape[0]) if vectors else None)

    def fake_batched(G, vectors=True, return_info=False):
        calls.append(("batched", tuple(G.shape), vectors))
        return torch.zeros(G.shape[:2]), (torch.eye(G.shape[1]).expand(G.shape[0], -1, -1) if vectors else None)

    monkeypatch.setattr(kernels, "syevj_dist", fake_dist)
    monkeypatch.setattr(kernels, "syevj", fake_one)
    monkeypatch.setattr(kernels, "syevj_batched", fake_batched)
    got = []
    done = lambda ev, U: got.append((ev.shape, None if U is None else U.shape))  # noqa: E731

    q = SolveQueue()
    sharded = _FakeDist(comm=1234)
    q.submit(torch.eye(6), done, dist=sharded)  # alone in its bucket, NCCL group: all ranks solve it together
    q.submit(torch.eye(4), done, vectors=False, dist=_FakeDist(comm=0))  # communicator not reachable: every rank alone
    q.submit(torch.eye(5), done)  # not sharded
    q.submit(torch.eye(3), done, dist=sharded)
    q.submit(torch.eye(3), done, dist=sharded)  # two of one shape: one batched call on every rank
    q.flush()
    assert ("dist", 1234, 2, (6, 6), True, True) in calls and sharded.arena_asked == 1
    assert ("one", (4, 4), False) in calls and ("one", (5, 5), True) in calls
    assert ("batched", (2, 3, 3), True) in calls
    assert len(calls) == 4 and len(got) == 5 and len(q) == 0
