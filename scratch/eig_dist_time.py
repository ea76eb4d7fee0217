"""Distributed two-level eigensolver (vvt_syevj_dist) against the one-GPU solve, under torchrun:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scratch/eig_dist_time.py [R ...]
ms by CUDA events, max over ranks; eigenvalue error against float64, orthogonality, residual."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import vivit_b200.kernels as k
from vivit_b200.dist import ShardedReduce
sr = ShardedReduce(dist.group.WORLD)
Rs = [int(a) for a in sys.argv[1:]] or [5120, 10240]
def ms_of(fn, reps=2):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), out
for R in Rs:
    torch.manual_seed(0)
    rank = int(0.9 * R)
    B = torch.randn(R, rank, dtype=torch.float64, device=dev) * torch.logspace(0, -3, rank, dtype=torch.float64, device=dev)
    G = (B @ B.t()).float(); del B
    dist.broadcast(G, 0)
    comm = sr.solver_comm(G)
    ms1, (ev1, U1, info1) = ms_of(lambda: k.syevj(G, True, return_info=True))
    del U1
    msn, (evn, Un, infon) = ms_of(lambda: k.syevj_dist(comm, sr.world, G, True, return_info=True))
    del Un
    p2p = sr.solver_arena(G)
    msd, (ev, U, info) = ms_of(lambda: k.syevj_dist(comm, sr.world, G, True, return_info=True, p2p=p2p))
    want = torch.linalg.eigvalsh(G.double())
    err = (ev.double() - want).abs().max().item() / want.abs().max().item()
    Ud = U.double()
    orth = (Ud.t() @ Ud - torch.eye(R, device=dev, dtype=torch.float64)).abs().max().item()
    resid = (G.double() @ Ud - Ud * ev.double()[None]).norm().item() / G.double().norm().item()
    if local == 0:
        print(f"R={R} world={sr.world}: one GPU {ms1:.1f} ms {info1} | NCCL hand-over {msn:.1f} ms {infon} | peer memory ({p2p}) {msd:.1f} ms {info} "
              f"evalerr={err:.2e} orth={orth:.2e} resid={resid:.2e}", flush=True)
    del U, Ud, want
dist.barrier(); dist.destroy_process_group()
