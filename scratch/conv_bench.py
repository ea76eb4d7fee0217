import sys, time; sys.path.insert(0, '.')
import torch, vivit_b200.kernels as k
torch.manual_seed(0)
dev = 'cuda'
V, N = 10, 128
shapes = {  # name: (Ci, Co, k, H_in, pad)
    'conv1': (3, 64, 5, 32, 0), 'conv2': (64, 96, 3, 14, 0), 'conv3': (96, 128, 3, 6, 1)}
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
def tm(f, n=reps):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, (ci, co, ks, h, pad) in shapes.items():
    if only and name != only: continue
    ho = h + 2 * pad - ks + 1
    S = torch.randn(V, N, co, ho, ho, device=dev)
    X = torch.rand(N, ci, h, h, device=dev)
    W = torch.randn(co, ci, ks, ks, device=dev)
    t_e = tm(lambda: k.v_emit_conv2d(S, X, (ks, ks), (1, 1), (pad, pad), (1, 1)))
    fl_e = 2 * V * N * co * ci * ks * ks * ho * ho
    msg = f"{name}: emit {t_e:.3f} ms ({fl_e / t_e / 1e9:.1f} TF/s, out {V*N*co*ci*ks*ks*4/1e6:.0f} MB)"
    if name != 'conv1':
        t_d = tm(lambda: k.sqrt_backprop_conv2d(S, W, (h, h), (1, 1), (pad, pad), (1, 1)))
        msg += f"  dgrad {t_d:.3f} ms ({fl_e / t_d / 1e9:.1f} TF/s)"
    print(msg, flush=True)
# pools
for (ch, hi) in [(64, 28), (96, 12), (128, 6)]:
    x = torch.rand(N, ch, hi, hi, device=dev)
    out, idx = torch.nn.functional.max_pool2d(x, 3, 2, 0, 1, True, return_indices=True)
    S = torch.randn(V, N, *out.shape[1:], device=dev)
    t = tm(lambda: k.sqrt_backprop_maxpool2d(S, idx, (hi, hi), (3, 3), (2, 2), (0, 0), (1, 1)))
    byts = (S.numel() + V * N * ch * hi * hi) * 4
    print(f"maxpool ch={ch} hi={hi}: {t:.3f} ms ({byts / t / 1e6:.0f} GB/s)")
