import sys, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
dev = 'cuda:0'
torch.manual_seed(0)
for R, D in [(128, 128), (128, 1024), (128, 8192), (128, 65536), (1280, 55296)]:
    V = torch.randn(R, D, device=dev)
    G = torch.zeros(R, R, device=dev)
    k.gram_dense_accum(G, V)
    want = V.double() @ V.double().t()
    d = (G.double() - want)
    diag_rel = (d.diag() / want.diag())
    off = d - torch.diag(d.diag())
    print(f"R={R} D={D}: max err/scale {d.abs().max().item()/want.abs().max().item():.2e}  diag rel err mean {diag_rel.mean().item():+.2e} (min {diag_rel.min().item():+.2e} max {diag_rel.max().item():+.2e})  offdiag max/scale {off.abs().max().item()/want.abs().max().item():.2e}", flush=True)
# positive matrix (all products positive): worst case for biased rounding
V = torch.rand(256, 16384, device=dev) + 0.5
G = torch.zeros(256, 256, device=dev); k.gram_dense_accum(G, V)
want = V.double() @ V.double().t()
print("positive V: rel err mean %+.2e max %.2e" % (((G.double()-want)/want).mean().item(), ((G.double()-want)/want).abs().max().item()))
ref32 = (V @ V.t()).double()
print("torch fp32 matmul same: rel err mean %+.2e max %.2e" % (((ref32-want)/want).mean().item(), ((ref32-want)/want).abs().max().item()))
