import sys, time, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
torch.manual_seed(0)
for R, rank, dt in [(2560, 2304, torch.float32), (5120, 4608, torch.float32), (1280, 1152, torch.float64)]:
    B = torch.randn(R, rank, dtype=torch.float64, device='cuda') * torch.logspace(0, -3, rank, dtype=torch.float64, device='cuda')
    G = (B @ B.t()).to(dt)
    k.syevj(G, True); torch.cuda.synchronize()
    t0 = time.time(); ev, U = k.syevj(G, True); torch.cuda.synchronize(); ms = (time.time() - t0) * 1e3
    want = torch.linalg.eigvalsh(G.double())
    print(f"R={R} {dt}: {ms:.1f} ms {k.last_syevj_info} evalerr={(ev.double()-want).abs().max().item()/want.abs().max().item():.2e}", flush=True)
