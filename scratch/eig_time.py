"""Eigensolver timings (ms, CUDA events) against torch.linalg.eigh (cuSOLVER) on the same GPU.
    python scratch/eig_time.py [R ...]"""
import sys, time, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
torch.manual_seed(0)
Rs = [int(a) for a in sys.argv[1:]] or [320, 1280, 2560, 5120]
def ms_of(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out
for R in Rs:
    rank = int(0.9 * R)
    B = torch.randn(R, rank, dtype=torch.float64, device='cuda') * torch.logspace(0, -3, rank, dtype=torch.float64, device='cuda')
    G = (B @ B.t()).float(); del B
    ms, (ev, U, info) = ms_of(lambda: k.syevj(G, True, return_info=True))
    ms_t, _ = ms_of(lambda: torch.linalg.eigh(G)) if R <= 10240 else (float('nan'), None)
    want = torch.linalg.eigvalsh(G.double())
    err = (ev.double() - want).abs().max().item() / want.abs().max().item()
    Ud = U.double()
    orth = (Ud.t() @ Ud - torch.eye(R, device='cuda', dtype=torch.float64)).abs().max().item()
    resid = (G.double() @ Ud - Ud * ev.double()[None]).norm().item() / G.double().norm().item()
    print(f"R={R}: {ms:.1f} ms (cuSOLVER {ms_t:.1f}) {info} evalerr={err:.2e} orth={orth:.2e} resid={resid:.2e}", flush=True)
    del U, Ud, want
