import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops
from kbench import timeit
dev = torch.device("cuda:0")
B, N, k, S, C = 128, 1028, 20, 7, 128
g = torch.Generator().manual_seed(0)
xyz = (torch.randn(B, N, 3, generator=g) * 0.05).to(dev)
idx = ops.knn3(xyz, xyz, k)[1]
dirn = torch.nn.functional.normalize(torch.randn(3, S * C, generator=g), dim=0).to(dev).requires_grad_()
out = ops.surface_conv(xyz, idx, dirn, S, C)
go = torch.randn_like(out)
print("surface fwd(argmax) ms", timeit(lambda: ops.surface_conv(xyz, idx, dirn, S, C)),
      "bwd ms", timeit(lambda: torch.autograd.grad(out, dirn, go, retain_graph=True)))
