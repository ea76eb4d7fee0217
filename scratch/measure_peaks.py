"""TF32 and FP64 dense GEMM throughput of this box through torch.matmul (cuBLAS), SURVEY 8d: the denominators
behind `frac_of_3xtf32_peak` (3xTF32 = TF32 peak / 3) and the fp64 Gram kernels."""
import json, sys, torch
torch.backends.cuda.matmul.allow_tf32 = True
out = {}
for name, dt, n in (("tf32", torch.float32, 8192), ("fp64", torch.float64, 4096)):
    a = torch.randn(n, n, device="cuda", dtype=dt); b = torch.randn(n, n, device="cuda", dtype=dt)
    for _ in range(3): a @ b
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_tflops"] = round(2 * n ** 3 / best / 1e9, 1)
    out[name + "_n"] = n
print(json.dumps(out))
