import torch, sys
sys.path.insert(0, '.')
import vivit_b200.kernels as k
torch.manual_seed(0)
dev = 'cuda:0'
for dtype in (torch.float32, torch.float64):
    for R, rank in [(33,33),(64,64),(64,20),(100,40),(100,100),(128,128),(320,288),(640,64)]:
        B = torch.randn(R, rank, dtype=torch.float64)
        G = (B @ B.t() / rank).to(dtype).to(dev)
        ev, U = k.syevj(G, True)
        want = torch.linalg.eigvalsh(G.double())
        err = (ev.double()-want).abs().max().item()/want.abs().max().item()
        tr = (ev.double().sum()-G.double().trace()).item()
        Ud = U.double()
        orth = (Ud.t()@Ud - torch.eye(R, device=dev, dtype=torch.float64)).abs().max().item()
        rec = (Ud @ torch.diag(ev.double()) @ Ud.t() - G.double()).abs().max().item()
        print(dtype, R, rank, 'sweeps', k.last_syevj_info, 'relerr %.2e trace_diff %.2e orth %.2e recon %.2e' % (err, tr, orth, rec))
