"""gram_linear_accum at bench shapes against float64 (with and without the bias fold)."""
import sys, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
torch.manual_seed(0)
for (C, N, n_out, n_in) in [(10, 512, 4096, 4096), (10, 1024, 2048, 2048), (10, 1024, 2048, 784), (10, 1024, 10, 2048), (10, 128, 512, 1152)]:
    S = (torch.randn(C, N, n_out, device='cuda') * 1e-2)
    Z = torch.rand(N, n_in, device='cuda')
    R = C * N
    Sd = S.double().reshape(R, n_out)
    SS = Sd @ Sd.t()
    P = (Z.double() @ Z.double().t())
    for bias in (False, True):
        want = (SS.reshape(C, N, C, N) * (P + (1.0 if bias else 0.0))[None, :, None, :]).reshape(R, R)
        G = torch.zeros(R, R, device='cuda')
        k.gram_linear_accum(G, S, Z, bias)
        err = (torch.triu(G).double() - torch.triu(want)).abs().max().item() / want.abs().max().item()
        print(f"C={C} N={N} out={n_out} in={n_in} bias={bias}: rel err (upper) {err:.2e}", flush=True)
    G = torch.zeros(R, R, device='cuda')
    k.gram_dense_accum(G, S.reshape(R, n_out))
    print(f"   dense S S^T: {(torch.triu(G).double() - torch.triu(SS)).abs().max().item() / SS.abs().max().item():.2e}", flush=True)
    del SS, P, want, G, Sd
