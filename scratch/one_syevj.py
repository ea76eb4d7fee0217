import sys, numpy as np, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
G = torch.from_numpy(np.load('scratch/G_c2.npy')).to('cuda:0')
if len(sys.argv) > 1 and sys.argv[1] == 'f64': G = G.double()
ev, U = k.syevj(G, True)
torch.cuda.synchronize()
print(k.last_syevj_info)
