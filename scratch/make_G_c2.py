"""Builds the cifar10_3c3d (N=128, C=10) Gram with the CPU oracle -> scratch/G_c2.npy (git-ignored)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch import nn
import bench
from oracle import reference_path as ref
w = bench.WORKLOADS["c2"]
model, x, y = bench.make_problem(w, torch.float32)
t0 = time.time()
sweep = ref.backward_sweep(model, nn.CrossEntropyLoss(), x, y, want_vivit=True)
gram = 0.0
for p in model.parameters():
    gram = gram + sweep.vivit[id(p)]["gram_mat"]()
G = ref.reshape_as_square(gram)
np.save("scratch/G_c2.npy", G.numpy())
print(G.shape, time.time() - t0)
