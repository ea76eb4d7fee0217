import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
dev = 'cuda:0'
G0 = torch.from_numpy(np.load('scratch/G_c2.npy'))
for dtype in (torch.float32, torch.float64):
    for name, G in [('c2', G0), ('c2[:320]', G0[:320, :320].contiguous()), ('rand640r64', None), ('c2[:32]', G0[:32,:32].contiguous())]:
        if G is None:
            B = torch.randn(640, 64, dtype=torch.float64); G = B @ B.t() / 64
        G = G.to(dtype).to(dev)
        R = G.shape[0]
        ev, U = k.syevj(G, True)
        torch.cuda.synchronize()
        t0 = time.time(); n = 3
        for _ in range(n): ev, U = k.syevj(G, True)
        torch.cuda.synchronize(); ms = (time.time() - t0) / n * 1e3
        t0 = time.time()
        for _ in range(n): torch.linalg.eigh(G)
        torch.cuda.synchronize(); ms_t = (time.time() - t0) / n * 1e3
        want = torch.linalg.eigvalsh(G.double())
        err = (ev.double() - want).abs().max().item() / want.abs().max().item()
        Ud = U.double()
        orth = (Ud.t() @ Ud - torch.eye(R, device=dev, dtype=torch.float64)).abs().max().item()
        resid = (G.double() @ Ud - Ud * ev.double()[None]).norm().item() / G.double().norm().item()
        print(f"{str(dtype):14s} {name:12s} R={R:5d} {ms:8.2f} ms (torch eigh {ms_t:7.2f} ms) info={k.last_syevj_info} evalerr={err:.2e} orth={orth:.2e} resid={resid:.2e}", flush=True)
