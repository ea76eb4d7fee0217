"""Round-2 starting point: the experimental wide stage of vvt_syevj (VVT_SYEVJ_WIDE=<inner sweeps>) against the
default path.  The switch is read once per process, so run it twice:

    python scratch/wide_check.py                    # default path
    VVT_SYEVJ_WIDE=1 VVT_SYEVJ_DEBUG=1 python scratch/wide_check.py   # wide stage, one inner sweep per visit
"""
import os, sys, time, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
torch.manual_seed(0)
print("VVT_SYEVJ_WIDE =", os.environ.get("VVT_SYEVJ_WIDE"))
for R in (1280, 2560, 5120):
    rank = int(0.9 * R)
    B = torch.randn(R, rank, dtype=torch.float64, device='cuda') * torch.logspace(0, -3, rank, dtype=torch.float64, device='cuda')
    G = (B @ B.t()).float()
    want = torch.linalg.eigvalsh(G.double())
    k.syevj(G, True); torch.cuda.synchronize()
    t0 = time.time(); ev, U = k.syevj(G, True); torch.cuda.synchronize(); ms = (time.time() - t0) * 1e3
    Ud = U.double()
    err = (ev.double() - want).abs().max().item() / want.abs().max().item()
    orth = (Ud.t() @ Ud - torch.eye(R, device='cuda', dtype=torch.float64)).abs().max().item()
    resid = (G.double() @ Ud - Ud * ev.double()[None]).norm().item() / G.double().norm().item()
    print(f"R={R}: {ms:.1f} ms {k.last_syevj_info} evalerr={err:.2e} orth={orth:.2e} resid={resid:.2e}", flush=True)
