"""Runs the dense Gram kernel alone (R=1280, D=110592: the conv3 weight of cifar10_3c3d at N=128)
so that `ncu --set full -k regex:<kernel>` can capture it.  Usage: python profiles/run_gram.py [R] [D] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vivit_b200 import kernels

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
D = int(sys.argv[2]) if len(sys.argv) > 2 else 110592
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
V = torch.randn(R, D, device="cuda")
G = torch.zeros(R, R, device="cuda")
for _ in range(reps):
    kernels.gram_dense_accum(G, V)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
kernels.gram_dense_accum(G, V)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"R={R} D={D}: {ms:.3f} ms, {R*(R+1)*D/ms/1e9:.1f} TFLOP/s (symmetric-aware algorithmic flops)")
