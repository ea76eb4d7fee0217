"""Host-side cost of the wrappers: per-call time of a tiny kernel, and a cProfile of one c2 step."""
import sys, time, cProfile, pstats, io, torch
sys.path.insert(0, '.')
import bench
from vivit_b200 import kernels
t = torch.ones(256, device='cuda')
for _ in range(100): kernels.scale_(t, 1.0)
torch.cuda.synchronize()
n = 5000
t0 = time.perf_counter()
for _ in range(n): kernels.scale_(t, 1.0)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"scale_ wrapper: {(t1 - t0) / n * 1e6:.1f} us per call (host)")
S = torch.randn(10, 128, 64, device='cuda'); W = torch.randn(64, 32, device='cuda')
for _ in range(100): kernels.sqrt_backprop_linear(S, W)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): kernels.sqrt_backprop_linear(S, W)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"sqrt_backprop_linear wrapper: {(t1 - t0) / n * 1e6:.1f} us per call (host)")
st = bench.Stepper(bench.WORKLOADS['c2'], torch.float32, torch.device('cuda:0'))
for _ in range(4): st.step_device()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): st.step_device()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
