import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, vivit_b200.kernels as k
G = torch.from_numpy(np.load('scratch/G_c2.npy')).cuda()
n = max(1, int(sys.argv[1])) if len(sys.argv) > 1 else 1
k.syevj(G, True); torch.cuda.synchronize()
t0 = time.time()
for _ in range(n): ev, U = k.syevj(G, True)
torch.cuda.synchronize()
print('ms', (time.time() - t0) / n * 1e3, k.last_syevj_info)
