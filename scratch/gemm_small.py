import sys; sys.path.insert(0, '.')
import torch, vivit_b200.kernels as k
def tm(f, n=20):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M, K in [(1216, 64), (1216, 128), (1216, 256), (1280, 1280), (512, 64), (128, 64)]:
    A = torch.randn(M, K, device='cuda'); C = torch.zeros(M, M, device='cuda')
    t0 = tm(lambda: k.gemm(A, A, out=C, alpha=-1.0, beta=1.0))
    t1 = tm(lambda: k.gemm(A, A, out=C, alpha=1.0, beta=0.0))
    t2 = tm(lambda: torch.matmul(A, A.t()))
    print(f"M=N={M} K={K}: vvt_gemm beta=1 {t0:.1f} us, beta=0 {t1:.1f} us, torch.matmul {t2:.1f} us")
