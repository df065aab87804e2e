import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops
from kbench import timeit
dev = torch.device("cuda:0")
B, N, k, S, C = 128, 1028, 20, 7, 128
g = torch.Generator().manual_seed(0)
xyz = (torch.randn(B, N, 3, generator=g) * 0.05).to(dev)
fm = torch.relu(torch.randn(B, N, C, generator=g)).to(dev)
rf = ops.knn_feat(fm, k)[1]
dirn = torch.nn.functional.normalize(torch.randn(3, S * C, generator=g), dim=0).to(dev)
P16 = torch.randn(B, N, (S + 1) * C, generator=g).to(dev).to(torch.bfloat16)
print(os.environ.get("HSP_GC2_VARIANT"), "tagged", timeit(lambda: ops._graph_conv_fwd_raw(xyz, rf, dirn, P16, S, C, True)),
      "nograd", timeit(lambda: ops._graph_conv_fwd_raw(xyz, rf, dirn, P16, S, C, False)))
