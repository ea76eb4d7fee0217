"""Worker of ``tests/test_dist_gpu.py`` (launched under torchrun, one rank per GPU, NCCL): the parameter-sharded
path against the unsharded one on the same GPU, with the partial-Gram exchange going through
``vvt_nccl_allreduce_gram`` (fused sub-sampling rescale)."""
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run(comp, model, x, y, groups):
    from vivit_b200 import backpack, extend

    model, loss_fn = extend(model), extend(nn.CrossEntropyLoss())
    with backpack(*comp.get_extensions(), extension_hook=comp.get_extension_hook(groups)):
        loss_fn(model(x), y).backward()
    for p in model.parameters():
        p.grad = None
    return [comp.get_result(g) for g in groups]


def check_distributed_solver(dev, sr):
    """``vvt_syevj_dist``: the rounds of the two-level eigensolver distributed over the ranks against the same
    solver on one GPU and against float64 LAPACK-style truth (``torch.linalg.eigvalsh`` in double).  The two-level
    path is forced on small matrices (``VVT_SYEVJ_WIDE_MIN``); 1088 columns give nine block pairs per round, an
    uneven split over two ranks."""
    from vivit_b200 import kernels

    os.environ["VVT_SYEVJ_WIDE_MIN"] = "512"
    try:
        comm = sr.solver_comm(torch.empty(2048, 2048, device=dev))
        assert comm != 0
        for R in (1024, 1088, 2048):
            gen = torch.Generator(device="cpu").manual_seed(R)
            rank = int(0.8 * R)
            B = torch.randn(R, rank, dtype=torch.float64, generator=gen) * torch.logspace(0, -3, rank, dtype=torch.float64)
            G = (B @ B.t()).float().to(dev)
            ev1, U1, info1 = kernels.syevj(G, True, return_info=True)
            evd, Ud, infod = kernels.syevj_dist(comm, sr.world, G, True, return_info=True)  # blocks by ncclSend / ncclRecv
            # ... and pushed into the next owner's factor through peer memory (CUDA IPC arenas): the same rotations
            # in the same order, so the same bits
            assert sr.solver_arena(G), "the peer-memory arena could not be mapped"
            evp, Up, infop = kernels.syevj_dist(comm, sr.world, G, True, return_info=True, p2p=True)
            assert infop == infod and torch.equal(evp, evd) and torch.equal(Up, Ud), (R, infop, infod)
            assert info1["converged"] and infod["converged"], (R, info1, infod)
            want = torch.linalg.eigvalsh(G.double())
            scale = want.abs().max()
            assert (evd.double() - want).abs().max() <= 1e-5 * scale, (R, (evd.double() - want).abs().max() / scale)
            assert (evd - ev1).abs().max() <= 1e-5 * scale
            Q = Ud.double()
            eye = torch.eye(R, dtype=torch.float64, device=dev)
            assert (Q.t() @ Q - eye).abs().max() < 5e-4, (R, (Q.t() @ Q - eye).abs().max())
            resid = (G.double() @ Q - Q * evd.double()[None]).norm() / G.double().norm()
            assert resid < 5e-5, (R, resid)
            # every rank holds the same eigenpairs, bit for bit
            both = [torch.empty_like(Ud) for _ in range(sr.world)]
            dist.all_gather(both, Ud.contiguous())
            assert all(torch.equal(both[0], t) for t in both), R
            # eigenvalues only
            ev0, none, _ = kernels.syevj_dist(comm, sr.world, G, False, return_info=True)
            assert none is None and (ev0 - evd).abs().max() <= 1e-6 * scale
    finally:
        del os.environ["VVT_SYEVJ_WIDE_MIN"]


def check_team_of_ranks(dev):
    """``dist.team_groups`` with more ranks than groups: the one group of this model gets a team of both ranks -- a
    sub-group of its own, not WORLD -- whose Gram (R = 640) is parameter-sharded inside the team and decomposed by
    the team together (two-level path forced on the small matrix; the peer-memory arena of this device belongs to
    WORLD already, so the blocks travel by ncclSend / ncclRecv on the team's communicator)."""
    import vivit_b200 as vv
    from vivit_b200.dist import team_groups

    torch.manual_seed(1)
    model = nn.Sequential(nn.Linear(20, 16), nn.ReLU(), nn.Linear(16, 10)).to(dev)
    x, y = torch.rand(64, 20, device=dev), torch.randint(0, 10, (64,), device=dev)
    top = lambda ev: list(range(ev.numel() - 4, ev.numel()))  # noqa: E731
    groups = [{"params": list(model.parameters()), "criterion": top}]
    own, team, teams = team_groups(groups, dist.group.WORLD)
    assert own == groups and team is not None and teams == [[0, 1]]
    ((ev1, vec1),) = run(vv.EighComputation(), model, x, y, groups)
    os.environ["VVT_SYEVJ_WIDE_MIN"] = "512"
    try:
        seen = []
        orig = vv.kernels.syevj_dist

        def spy(*a, **k):
            seen.append(k.get("p2p"))
            return orig(*a, **k)

        vv.kernels.syevj_dist = spy
        ((evs, vecs),) = run(vv.EighComputation(process_group=team, gather=True), model, x, y, groups)
        vv.kernels.syevj_dist = orig
        assert seen == [False], seen  # the team's solve was distributed, over NCCL
    finally:
        del os.environ["VVT_SYEVJ_WIDE_MIN"]
    assert torch.allclose(ev1, evs, rtol=1e-4, atol=1e-7), (ev1, evs)
    a = torch.cat([v.flatten(1) for v in vec1], 1).double()
    b = torch.cat([v.flatten(1) for v in vecs], 1).double()
    assert (a @ b.t()).abs().diag().min() > 1 - 1e-3


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
    import vivit_b200 as vv
    from vivit_b200.dist import ShardedReduce

    torch.manual_seed(0)
    model = nn.Sequential(
        nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
        nn.Linear(8 * 4 * 4, 16), nn.ReLU(), nn.Linear(16, 5),
    ).to(dev)
    x, y = torch.rand(6, 3, 8, 8, device=dev), torch.randint(0, 5, (6,), device=dev)
    top = lambda ev: list(range(ev.numel() - 3, ev.numel()))  # noqa: E731
    groups = [{"params": list(model.parameters()), "criterion": top}]

    sr = ShardedReduce(pg)
    assert sr._nccl_comm(x) != 0, "the NCCL communicator is not reachable: the fused exchange would not run"

    for sub in (None, [4, 1, 0]):
        ((ev1, vec1),) = run(vv.EighComputation(subsampling=sub), model, x, y, groups)
        ((evs, vecs),) = run(vv.EighComputation(subsampling=sub, process_group=pg, gather=True), model, x, y, groups)
        assert torch.allclose(ev1, evs, rtol=1e-4, atol=1e-7), (sub, ev1, evs)
        a = torch.cat([v.flatten(1) for v in vec1], 1).double()
        b = torch.cat([v.flatten(1) for v in vecs], 1).double()
        assert (a @ b.t()).abs().diag().min() > 1 - 1e-3, sub
        ((g1, l1),) = run(vv.DirectionalDerivativesComputation(subsampling_ggn=sub), model, x, y, groups)
        ((gs, ls),) = run(vv.DirectionalDerivativesComputation(subsampling_ggn=sub, process_group=pg), model, x, y, groups)
        assert torch.allclose(g1.abs(), gs.abs(), rtol=1e-3, atol=1e-5) and torch.allclose(l1, ls, rtol=1e-3, atol=1e-6), sub
    # every rank holds the same result (replicated Gram-space work is deterministic)
    gathered = [torch.empty_like(evs) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, evs.contiguous())
    assert all(torch.equal(gathered[0], t) for t in gathered)
    check_distributed_solver(dev, sr)
    check_team_of_ranks(dev)
    dist.barrier()
    dist.destroy_process_group()
    if local == 0:
        print("dist gpu worker ok")


if __name__ == "__main__":
    main()
