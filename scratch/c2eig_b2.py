"""Two cifar10_3c3d Grams in one batched solve against two single solves."""
import sys, torch
sys.path.insert(0, '.')
import bench, vivit_b200 as vv
from vivit_b200 import kernels
w = bench.WORKLOADS['c2']
st = bench.Stepper(w, torch.float32, torch.device('cuda:0'))
grabbed = []; orig = kernels.syevj
def spy(G, vectors=True, **kw):
    grabbed.append(G.clone()); return orig(G, vectors, **kw)
kernels.syevj = spy
st._pass(vv.EighComputation(), st.x, st.y)
kernels.syevj = orig
G = grabbed[0]
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
for B in (1, 2, 3, 4):
    GG = torch.stack([G] * B).contiguous()
    ms, out = timeit(lambda: kernels.syevj_batched(GG, True, return_info=True))
    print('B', B, 'ms', ms, out[2])
