import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
dev = 'cuda:0'
G0 = torch.from_numpy(np.load('scratch/G_c2.npy'))
torch.manual_seed(0)
B = torch.randn(5120, 4096, dtype=torch.float32, device=dev); G5 = (B @ B.t() / 4096).cpu(); del B
for dtype in (torch.float32, torch.float64):
    for name, G in [('c2', G0), ('c2[:320]', G0[:320, :320].contiguous()), ('c2[:640]', G0[:640, :640].contiguous()), ('rand5120r4096', G5)]:
        if dtype == torch.float64 and G.shape[0] > 2000: continue
        G = G.to(dtype).to(dev)
        R = G.shape[0]
        for vectors in (True, False):
            ev, U = k.syevj(G, vectors); torch.cuda.synchronize()
            n = 3 if R < 2000 else 1
            t0 = time.time()
            for _ in range(n): ev, U = k.syevj(G, vectors)
            torch.cuda.synchronize(); ms = (time.time() - t0) / n * 1e3
            t0 = time.time()
            for _ in range(n): (torch.linalg.eigh(G) if vectors else torch.linalg.eigvalsh(G))
            torch.cuda.synchronize(); ms_t = (time.time() - t0) / n * 1e3
            want = torch.linalg.eigvalsh(G.double())
            err = (ev.double() - want).abs().max().item() / want.abs().max().item()
            msg = f"{str(dtype):14s} {name:14s} R={R:5d} vec={int(vectors)} {ms:8.2f} ms (torch {ms_t:7.2f} ms) info={k.last_syevj_info} evalerr={err:.2e}"
            if vectors:
                Ud = U.double()
                orth = (Ud.t() @ Ud - torch.eye(R, device=dev, dtype=torch.float64)).abs().max().item()
                resid = (G.double() @ Ud - Ud * ev.double()[None]).norm().item() / G.double().norm().item()
                msg += f" orth={orth:.2e} resid={resid:.2e}"
            print(msg, flush=True)
