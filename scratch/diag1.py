"""GPU diagnostics: (1) the failing gamma case, (2) dump the c2 Gram, (3) syevj phase timing."""
import copy, os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from torch import nn
import vivit_b200 as vv
import vivit_b200.kernels as k
from oracle import reference_path as ref
from tests.problems import PROBLEMS, IDS, GROUPINGS, GROUPING_IDS, make_top_k, constant_damping
dev = 'cuda:0'
captured = []
orig = k.syevj
def spy(G, vectors=True):
    ev, U = orig(G, vectors)
    captured.append((G.clone(), ev.clone(), None if U is None else U.clone()))
    return ev, U
k.syevj = spy
import vivit_b200.optim.directional_derivatives as dd
p = PROBLEMS[IDS.index('allcnnc-mini')]
grouping = GROUPINGS[GROUPING_IDS.index('weights_and_biases')]
cm, cl, cx, cy = p.make(torch.float32, 'cpu')
gm = copy.deepcopy(cm).to(dev)
crit = make_top_k(1, must_exceed=1e-4)
cg = grouping(cm, criterion=crit, damping=constant_damping(1.0))
table = {id(a): b for a, b in zip(cm.parameters(), gm.parameters())}
gg = [{**g, 'params': [table[id(q)] for q in g['params']]} for g in cg]
comp = vv.DirectionalDerivativesComputation()
m, lf = vv.extend(gm), vv.extend(copy.deepcopy(cl).to(dev))
with vv.backpack(*comp.get_extensions(), extension_hook=comp.get_extension_hook(gg)):
    lf(m(cx.to(dev)), cy.to(dev)).backward()
want = ref.directional_derivatives(cm, cl, cx, cy, cg, None, None)
for g, (wg, wl) in zip(gg, want):
    gam, lam = comp.get_result(g)
    print('gamma relerr', ((gam.abs().cpu() - wg.abs()).abs().max() / wg.abs().max()).item(), 'lam relerr', ((lam.cpu()-wl).abs().max()/wl.abs().max()).item())
for G, ev, U in captured:
    Gd = G.double()
    w, Q = torch.linalg.eigh(Gd)
    print('R', G.shape[0], 'top evals ours', ev[-3:].tolist(), 'fp64', w[-3:].tolist())
    u, q = U[:, -1].double(), Q[:, -1]
    print('  top evec err', min((u - q).abs().max().item(), (u + q).abs().max().item()), 'resid', ((Gd @ u - ev[-1].double() * u).norm() / w[-1]).item(), 'norm', u.norm().item(), 'info', k.last_syevj_info)
    # alignment of the errors with other eigenvectors
    coef = Q.t() @ u
    top = coef.abs().argsort(descending=True)[:4]
    print('  coef on fp64 evecs', [(int(i), float(coef[i]), float(w[i])) for i in top])
k.syevj = orig

# (2) c2 Gram
sys.path.insert(0, '.')
import bench
captured.clear(); k.syevj = spy
w = bench.WORKLOADS['c2']
st = bench.Stepper(w, torch.float32, torch.device(dev))
import vivit_b200.linalg.eigh as eh
st._pass(vv.EighComputation(), st.x, st.y)
k.syevj = orig
G = captured[0][0]
os.makedirs('gpurun_out', exist_ok=True)
np.save('gpurun_out/G_c2.npy', G.cpu().numpy())
print('saved c2 Gram', G.shape)

# (3) syevj timing
os.environ['VVT_SYEVJ_DEBUG'] = '1'
ev, U = k.syevj(G, True)
del os.environ['VVT_SYEVJ_DEBUG']
for name, GG in [('c2', G), ('c2[:320]', G[:320, :320].contiguous()), ('c2[:640]', G[:640, :640].contiguous())]:
    for vectors in (True, False):
        k.syevj(GG, vectors); torch.cuda.synchronize()
        t0 = time.time(); n = 3
        for _ in range(n): ev, U = k.syevj(GG, vectors)
        torch.cuda.synchronize(); ms = (time.time() - t0) / n * 1e3
        t0 = time.time()
        for _ in range(n): (torch.linalg.eigh(GG) if vectors else torch.linalg.eigvalsh(GG))
        torch.cuda.synchronize(); ms_t = (time.time() - t0) / n * 1e3
        want = torch.linalg.eigvalsh(GG.double())
        err = (ev.double() - want).abs().max().item() / want.abs().max().item()
        print(f"{name:10s} vectors={vectors} {ms:8.2f} ms (torch {ms_t:7.2f} ms) info={k.last_syevj_info} evalerr={err:.2e}", flush=True)
