import sys; sys.path.insert(0, '.')
import torch, vivit_b200.kernels as k
dev = 'cuda'
def tm(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
R = 1280
for D in (4800, 55296, 110592):
    V = torch.randn(R, D, device=dev)
    for K in (1, 10, 16):
        U = torch.randn(K, R, device=dev)
        n2 = torch.zeros(K, dtype=torch.float64, device=dev)
        t = tm(lambda: k.backtransform_dense(U, V, n2))
        print(f"backtransform R={R} D={D} K={K}: {t:.3f} ms  {(R*D+K*R+K*D)*4/t/1e6:.0f} GB/s")
    v = torch.randn(R, device=dev)
    t = tm(lambda: k.v_apply_dense(v, V))
    print(f"v_apply R={R} D={D}: {t:.3f} ms  {(R*D+R+D)*4/t/1e6:.0f} GB/s")
