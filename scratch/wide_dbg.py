"""Diagnostics of one wide round (vvt_dbg_wide_round): which interpretation of the operands does H match?"""
import ctypes, math, os, sys, torch
sys.path.insert(0, '.')
from vivit_b200 import _lib
lib = _lib.load()
fn = lib.vvt_dbg_wide_round; fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
Np = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
L0 = (torch.randn(Np, Np, dtype=torch.float64) / math.sqrt(Np)).float().cuda()
L = L0.clone()
pairs = Np // 128
H = torch.zeros(pairs, 128, 128, device='cuda'); Qt = torch.zeros_like(H); flag = torch.zeros(pairs, dtype=torch.int32, device='cuda')
st = fn(L.data_ptr(), H.data_ptr(), Qt.data_ptr(), flag.data_ptr(), Np, -1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print('variant', os.environ.get('VVT_WIDE_DESC'), 'status', st, 'flag', flag.tolist())
P = L0[:, :128].double()
want = P.t() @ P
h = H[0].double()
print('H absmax', h.abs().max().item(), 'zeros frac', (h == 0).float().mean().item(), 'nan', torch.isnan(h).any().item())
print('err vs P^T P', (h - want).abs().max().item(), 'scale', want.abs().max().item())
alt = P[:128] @ P[:128].t()
print('err vs P[:128] P[:128]^T', (h - alt).abs().max().item())
print('diag H', h.diagonal()[:6].tolist(), 'want', want.diagonal()[:6].tolist())
print('H[0,:6]', h[0, :6].tolist(), 'want', want[0, :6].tolist())
if flag[0].item():
    Q = Qt[0].double().t()
    print('Q orth', (Q.t() @ Q - torch.eye(128, dtype=torch.float64, device='cuda')).abs().max().item())
    Pn = L[:, :128].double()
    print('apply err', (Pn - P @ Q).abs().max().item())
    Hn = Pn.t() @ Pn
    print('offdiag intra before/after', (want[:64, :64] - torch.diag(want.diagonal()[:64])).abs().max().item(), (Hn[:64, :64] - torch.diag(Hn.diagonal()[:64])).abs().max().item())
