#!/usr/bin/env python
"""Launch each hot kernel a few times at bench shapes (for `ncu -k regex:...`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops
dev = torch.device("cuda:0")
B, N, k, S, C = 128, 1028, 20, 7, 128
g = torch.Generator().manual_seed(0)
xyz = (torch.randn(B, N, 3, generator=g) * 0.05).to(dev)
fm = torch.relu(torch.randn(B, N, C, generator=g)).to(dev)
which = sys.argv[1:] or ["knn3", "knn_feat", "orl", "surface", "graph"]
for it in range(2):
    idx = ops.knn3(xyz, xyz, k)[1]
    if "knn_feat" in which:
        rf = ops.knn_feat(fm, k)[1]
    if "orl" in which:
        ops.orl_global(fm, idx)
        ops.orl_global(fm.clone().requires_grad_(), idx)
    if "surface" in which:
        dirn = torch.nn.functional.normalize(torch.randn(3, S * C, generator=g), dim=0).to(dev)
        ops.surface_conv(xyz, idx, dirn, S, C)
        ops.surface_conv(xyz, idx, dirn.clone().requires_grad_(), S, C)
torch.cuda.synchronize()
