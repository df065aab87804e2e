import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops
from kbench import timeit
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (B, N, D, k) in ((128, 1028, 128, 20), (128, 257, 128, 20), (128, 257, 256, 20), (128, 64, 256, 8)):
    fm = torch.relu(torch.randn(B, N, D, generator=g)).to(dev)
    print(os.environ.get("HSP_KF_NOSEL"), B, N, D, k, "ms", round(timeit(lambda: ops.knn_feat(fm, k)), 4))
