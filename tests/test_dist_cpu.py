"""Parameter-sharded path (``vivit_b200/dist.py``) with two ``gloo`` processes on CPU.

Each rank owns a dim-0 slice of every parameter, assembles a partial Gram and the partial
Grams are summed with one all-reduce; the results must equal the single-process oracle.
The kernel layer is the plain-torch test double (there is no GPU here); everything above
it -- extensions, factors, hooks, the four Computations, ``ShardedReduce`` -- is shipped code.
"""

import os
import socket
import sys
import traceback

import pytest
import torch
import torch.multiprocessing as mp
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    torch.manual_seed(0)
    model = nn.Sequential(
        nn.Conv2d(2, 5, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
        nn.Linear(5 * 3 * 3, 7), nn.Sigmoid(), nn.Linear(7, 4),
    ).double()
    x = torch.rand(5, 2, 6, 6, dtype=torch.float64)
    y = torch.randint(0, 4, (5,))
    return model, x, y


def _top(k):
    return lambda ev: list(range(max(0, ev.numel() - k), ev.numel()))


def _damping(evals, evecs, gammas, lambdas):
    return torch.ones_like(evals)


def _groups(model, grouping):
    extra = {"criterion": _top(3), "damping": _damping}
    if grouping == "one":
        return [{"params": list(model.parameters()), **extra}]
    return [{"params": list(m.parameters()), **extra} for m in model if list(m.parameters())]


class _MonkeyPatch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _worker(rank, world, port, grouping, errors):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import torch.distributed as dist

        import tests._torch_kernels as double
        from oracle import reference_path as ref

        double.install(_MonkeyPatch())
        dist.init_process_group(
            "gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world
        )
        import vivit_b200 as vv
        from vivit_b200.dist import shard_bounds

        loss_fn = nn.CrossEntropyLoss()

        def run(comp):
            model, x, y = _problem()
            groups = _groups(model, grouping)
            m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
            with vv.backpack(*comp.get_extensions(), extension_hook=comp.get_extension_hook(groups)):
                lf(m(x), y).backward()
            return [comp.get_result(g) for g in groups], groups

        def own(t, p_rows):  # this rank's dim-0 slice of a parameter-shaped oracle tensor
            lo, hi = shard_bounds(p_rows, rank, world)
            return t[lo:hi]

        model0, x0, y0 = _problem()
        groups0 = _groups(model0, grouping)
        pg = dist.group.WORLD

        # eigenvalues: identical on every rank, equal to the unsharded oracle
        res, _ = run(vv.EigvalshComputation(process_group=pg))
        want = ref.eigvalsh(model0, loss_fn, x0, y0, groups0)
        for got, w in zip(res, want):
            assert torch.allclose(got, w, rtol=1e-9, atol=1e-12), (got - w).abs().max()

        # eigenvectors: sharded along dim 0 (and gathered on request)
        want = ref.eigh(model0, loss_fn, x0, y0, groups0)
        for gather in (False, True):
            res, groups = run(vv.EighComputation(process_group=pg, gather=gather))
            for (ev, evecs), (wev, wvecs), g in zip(res, want, groups):
                assert torch.allclose(ev, wev, rtol=1e-9, atol=1e-12)
                for e, w_, p in zip(evecs, wvecs, g["params"]):
                    w_ = w_ if gather else torch.stack([own(v, p.shape[0]) for v in w_])
                    assert e.shape == w_.shape, (e.shape, w_.shape)
                    sign = torch.ones(e.shape[0], dtype=e.dtype)
                    # fix the sign per direction with the full (all-rank) inner product
                    dots = torch.stack([(a * b).sum() for a, b in zip(e, w_)])
                    if not gather:
                        dist.all_reduce(dots)
                    sign = torch.sign(dots)
                    for k in range(e.shape[0]):
                        assert torch.allclose(e[k] * sign[k], w_[k], rtol=1e-6, atol=1e-9)

        # directional derivatives live in Gram space: replicated on every rank
        res, _ = run(vv.DirectionalDerivativesComputation(process_group=pg))
        want = ref.directional_derivatives(model0, loss_fn, x0, y0, groups0)
        for (gam, lam), (wg, wl) in zip(res, want):
            assert torch.allclose(gam.abs(), wg.abs(), rtol=1e-7, atol=1e-10)
            assert torch.allclose(lam, wl, rtol=1e-7, atol=1e-10)

        # Newton step: sharded along dim 0
        res, groups = run(vv.DirectionalDampedNewtonComputation(process_group=pg))
        want = ref.directional_damped_newton(model0, loss_fn, x0, y0, groups0)
        for steps, wsteps, g in zip(res, want, groups):
            for s, w_, p in zip(steps, wsteps, g["params"]):
                w_ = own(w_, p.shape[0])
                assert s.shape == w_.shape
                assert torch.allclose(s, w_, rtol=1e-7, atol=1e-10), (s - w_).abs().max()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        errors.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("grouping", ["one", "layer"])
def test_two_rank_parameter_sharding_matches_oracle(grouping):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grouping, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("a rank hung")
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    assert not msgs, "\n".join(f"rank {r}:\n{t}" for r, t in msgs)
    assert all(p.exitcode == 0 for p in procs)


def test_shard_bounds_and_group_assignment():
    from vivit_b200.dist import assign_groups, shard_bounds

    for size in (1, 7, 10, 64, 100):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(size, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == size
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            widths = [hi - lo for lo, hi in cuts]
            assert max(widths) - min(widths) <= 1
    owner = assign_groups([5.0, 1.0, 1.0, 1.0, 4.0], 2)
    loads = [sum(c for c, o in zip([5.0, 1.0, 1.0, 1.0, 4.0], owner) if o == r) for r in range(2)]
    assert sorted(loads) == [6.0, 6.0]


def _worker_groups(rank, world, port, errors):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import torch.distributed as dist

        import tests._torch_kernels as double
        from oracle import reference_path as ref

        double.install(_MonkeyPatch())
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        import vivit_b200 as vv
        from vivit_b200.dist import local_groups

        model, x, y = _problem()
        groups = _groups(model, "layer")
        own, owner = local_groups(groups, dist.group.WORLD)
        assert sorted(set(owner)) == list(range(world)) and len(owner) == len(groups)
        assert [g for g, o in zip(groups, owner) if o == rank] == own
        # no more ranks than groups: team_groups is local_groups (teams of one, no sub-group)
        from vivit_b200.dist import team_groups

        own_t, team, teams = team_groups(groups, dist.group.WORLD)
        assert team is None and teams == [[o] for o in owner] and len(own_t) == len(own)
        assert all(a is b for a, b in zip(own_t, own))
        comp = vv.EighComputation()  # no process group: whole groups, no collective
        m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
        with vv.backpack(*comp.get_extensions(), extension_hook=comp.get_extension_hook(own)):
            lf(m(x), y).backward()
        model0, x0, y0 = _problem()
        want = ref.eigh(model0, nn.CrossEntropyLoss(), x0, y0, _groups(model0, "layer"))
        for g, (wev, wvecs), o in zip(groups, want, owner):
            if o != rank:
                with pytest.raises(KeyError):
                    comp.get_result(g)
                continue
            ev, evecs = comp.get_result(g)
            assert torch.allclose(ev, wev, rtol=1e-9, atol=1e-12)
            for e, w_ in zip(evecs, wvecs):
                for k in range(e.shape[0]):
                    sign = torch.sign((e[k] * w_[k]).sum())
                    assert torch.allclose(e[k] * sign, w_[k], rtol=1e-6, atol=1e-9)
        # every group is solved by exactly one rank
        mine = torch.tensor([float(o == rank) for o in owner])
        dist.all_reduce(mine)
        assert torch.equal(mine, torch.ones(len(groups)))
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        errors.put((rank, traceback.format_exc()))


def test_two_rank_block_diagonal_groups_need_no_collective():
    """block-diagonal groups are assigned to ranks whole (``dist.local_groups``)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    procs = [ctx.Process(target=_worker_groups, args=(r, world, port, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("a rank hung")
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    assert not msgs, "\n".join(f"rank {r}:\n{t}" for r, t in msgs)
    assert all(p.exitcode == 0 for p in procs)


def test_partial_grams_of_all_shards_sum_to_the_full_gram_even_with_empty_shards():
    """More ranks than output channels: some ranks own an empty slice of a parameter and contribute nothing.
    The shards of all eight 'ranks' are evaluated one after the other in this process (test double)."""
    import tests._torch_kernels as double
    from oracle import reference_path as ref

    double.install(_MonkeyPatch())
    import vivit_b200 as vv
    from vivit_b200.extensions.hooks import GramSqrtGGNExact

    world, total = 8, None
    for rank in range(world):
        model, x, y = _problem()  # last layer has 4 outputs, the conv layer 5 channels
        ext = vv.SqrtGGNExact(lazy=True)
        ext._shard = (rank, world)
        hook = GramSqrtGGNExact()
        m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
        with vv.backpack(ext, extension_hook=hook):
            lf(m(x), y).backward()
        total = hook.get_result() if total is None else total + hook.get_result()
    model, x, y = _problem()
    want, _ = ref.gram_sqrt_ggn(model, nn.CrossEntropyLoss(), x, y)
    assert torch.allclose(total, want, rtol=1e-9, atol=1e-12), (total - want).abs().max()


def _worker_teams(rank, world, port, errors):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import torch.distributed as dist

        import tests._torch_kernels as double
        from oracle import reference_path as ref

        double.install(_MonkeyPatch())
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        import vivit_b200 as vv
        from vivit_b200.dist import team_groups

        model, x, y = _problem()
        groups = _groups(model, "layer")  # three groups, four ranks: one team of two, two teams of one
        own, team, teams = team_groups(groups, dist.group.WORLD)
        assert sorted(r for t in teams for r in t) == list(range(world)) and len(teams) == len(groups)
        assert sorted(len(t) for t in teams) == [1, 1, 2]
        assert len(own) == 1
        which = next(i for i, g in enumerate(groups) if g is own[0])
        assert rank in teams[which] and (team is not None) == (len(teams[which]) > 1)
        kw = {"process_group": team, "gather": True} if team is not None else {}
        comp = vv.EighComputation(**kw)
        m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
        with vv.backpack(*comp.get_extensions(), extension_hook=comp.get_extension_hook(own)):
            lf(m(x), y).backward()
        model0, x0, y0 = _problem()
        want = ref.eigh(model0, nn.CrossEntropyLoss(), x0, y0, _groups(model0, "layer"))
        (wev, wvecs) = want[which]
        ev, evecs = comp.get_result(own[0])
        assert torch.allclose(ev, wev, rtol=1e-9, atol=1e-12)
        for e, w_ in zip(evecs, wvecs):
            assert e.shape == w_.shape  # gathered inside the team
            for k in range(e.shape[0]):
                sign = torch.sign((e[k] * w_[k]).sum())
                assert torch.allclose(e[k] * sign, w_[k], rtol=1e-6, atol=1e-9)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        errors.put((rank, traceback.format_exc()))


def test_more_ranks_than_groups_work_in_teams():
    """``dist.team_groups``: four ranks, three block-diagonal groups -- the costliest group gets a team of two ranks
    (parameter-sharded inside the team), the others one rank each; no collective between teams."""
    from vivit_b200.dist import team_sizes

    assert team_sizes([1.0, 1.0, 1.0, 1.0], 8) == [2, 2, 2, 2]
    assert team_sizes([3.0, 1.0], 4) == [3, 1] and sum(team_sizes([1.5, 1.2, 1.1], 8)) == 8
    world, port = 4, _free_port()
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    procs = [ctx.Process(target=_worker_teams, args=(r, world, port, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("a rank hung")
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    assert not msgs, "\n".join(f"rank {r}:\n{t}" for r, t in msgs)
    assert all(p.exitcode == 0 for p in procs)
