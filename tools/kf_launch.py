import os, sys, torch
sys.path.insert(0, "/root/repo")
import hspose_b200.ops as ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fm = torch.relu(torch.randn(B, 1028, 128, generator=g)).to(dev)
ops.knn_feat(fm, 20); torch.cuda.synchronize()
ops.knn_feat(fm, 20); torch.cuda.synchronize()
