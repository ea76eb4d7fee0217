"""Eigensolver on the cifar10_3c3d Gram (bench c2): ms for the settings given in the environment."""
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
for _ in range(2): orig(G, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ev, U, info = orig(G, True, return_info=True)
e1.record(); torch.cuda.synchronize()
want = torch.linalg.eigvalsh(G.double())
Ud = U.double()
print('ms', e0.elapsed_time(e1) / 5, info, 'evalerr', ((ev.double() - want).abs().max() / want.abs().max()).item(),
      'resid', ((G.double() @ Ud - Ud * ev.double()[None]).norm() / G.double().norm()).item(),
      'orth', (Ud.t() @ Ud - torch.eye(G.shape[0], device='cuda', dtype=torch.float64)).abs().max().item())
