"""Runs the Criteo-shape dense head a few times (for ncu): [65536, 3328] -> 1024 -> 512 -> 256 -> 1."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hugectr_backend_b200 as hb

torch.cuda.set_device(0)
dims = [3328, 1024, 512, 256, 1]
rng = np.random.default_rng(0)
w = [(rng.standard_normal((dims[l + 1], dims[l])) / np.sqrt(dims[l])).astype(np.float32) for l in range(4)]
mlp = hb.DenseMlp(0, w, [np.zeros(d, np.float32) for d in dims[1:]], [1, 1, 1, 0])
x = torch.randn((65536, dims[0]), device="cuda")
y = torch.empty((65536, 1), device="cuda")
for _ in range(4):
    mlp.forward(x, 65536, y)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
