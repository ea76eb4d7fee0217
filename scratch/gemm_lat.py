import sys, time, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
torch.manual_seed(0)
def bench(M, K, n=200):
    A = torch.randn(M, K, device='cuda')
    C = torch.zeros(M, M, device='cuda')
    for _ in range(5): k.gemm(A, A, out=C, alpha=-1e-3, beta=1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(n): k.gemm(A, A, out=C, alpha=-1e-3, beta=1.0)
    e1.record(); torch.cuda.synchronize()
    print(f"gemm M=N={M} K={K}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/call on device ({(time.time() - t0) / n * 1e6:.1f} us wall)", flush=True)
for M in (1216, 640, 256, 64):
    bench(M, 64)
bench(1280, 1280, 50)
# whole solver, and the sweeps alone via info
import numpy as np
B = torch.randn(1280, 1152, dtype=torch.float64, device='cuda') * torch.logspace(0, -3, 1152, dtype=torch.float64, device='cuda')
G = (B @ B.t()).float()
k.syevj(G, True); torch.cuda.synchronize()
t0 = time.time(); k.syevj(G, True); torch.cuda.synchronize(); print('syevj 1280 ms', (time.time() - t0) * 1e3, k.last_syevj_info)
