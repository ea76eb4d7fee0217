import sys, time, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
dev = 'cuda:0'
torch.manual_seed(0)
def check(R, D, sym=True, N=None):
    V = torch.randn(R, D, device=dev)
    if sym:
        G = torch.zeros(R, R, device=dev)
        k.gram_dense_accum(G, V)
        want = V.double() @ V.double().t()
    else:
        g = torch.randn(N, D, device=dev)
        G = torch.zeros(R, N, device=dev)
        k.gram_cross_accum(G, V, g)
        want = V.double() @ g.double().t()
    torch.cuda.synchronize()
    err = (G.double() - want).abs().max().item() / want.abs().max().item()
    print(f"R={R} D={D} sym={sym} N={N}: rel-to-scale err {err:.2e}", flush=True)
    return err
for R, D in [(128, 128), (128, 256), (256, 4096), (200, 1000), (1280, 4800), (1280, 55296)]:
    check(R, D)
check(1280, 4800, sym=False, N=128)
check(300, 5000, sym=False, N=77)
# timing
R, D = 1280, 110592
V = torch.randn(R, D, device=dev); G = torch.zeros(R, R, device=dev)
for _ in range(3): k.gram_dense_accum(G, V)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): k.gram_dense_accum(G, V)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"R={R} D={D}: {ms:.3f} ms, {R*(R+1)*D/ms/1e9:.1f} TFLOP/s (symmetric-aware)")
