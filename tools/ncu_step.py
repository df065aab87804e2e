#!/usr/bin/env python
"""Runs warm-up steps, then exactly ONE bench train step inside a cudaProfiler range
(use with `ncu --profile-from-start off ...`).  Never a source of bench numbers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hspose_b200 import parallel  # noqa: E402
from hspose_b200.HSPose import HSPose  # noqa: E402
from hspose_b200.synth import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
amp = (sys.argv[2] if len(sys.argv) > 2 else "bf16") == "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = HSPose("PoseNet_only", chamfer_w=1.0).to(dev).train()
flat = parallel.FlatGradients(model.posenet.parameters())
opt = torch.optim.Adam(flat.params, lr=1e-4, fused=True)
batch = {k: v.to(dev) for k, v in synth_batch(B, 1028, seed=1, train=True).items()}


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        _, losses = model(**batch, do_loss=True)
    total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
    flat.zero()
    total.backward()
    flat.clip_(5.0)
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
