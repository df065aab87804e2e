#!/usr/bin/env python
"""Runs warm-up steps, then exactly ONE bench train step inside a cudaProfiler range
(use with `ncu --profile-from-start off ...`).  Never a source of bench numbers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hspose_b200.engine import TrainStep  # noqa: E402
from hspose_b200.HSPose import HSPose  # noqa: E402
from hspose_b200.synth import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
amp = (sys.argv[2] if len(sys.argv) > 2 else "bf16") == "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = HSPose("PoseNet_only", chamfer_w=1.0).to(dev).train()
trainer = TrainStep(model, lr=1e-4, clip=5.0, amp=amp, graph=False, optimizer=(sys.argv[3] if len(sys.argv) > 3 else "adam"))
batch = {k: v.to(dev) for k, v in synth_batch(B, 1028, seed=1, train=True).items()}

for _ in range(2):
    trainer(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
trainer(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
