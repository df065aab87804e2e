#!/usr/bin/env python
"""Debug: per-parameter cosine between fp32-path and bf16-path gradients (RF-F tables forced)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hspose_b200.synth import fill_params, synth_batch
import hspose_b200.flags as hf
from hspose_b200 import gcn3d
from hspose_b200.HSPose import HSPose
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cuda = torch.device("cuda:0")
F = hf.get_flags()
for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro"):
    setattr(F, n, 0.0)
rf, res = [], {}
for mode in ("fp32", "bf16"):
    F.train, F.gcn_n_num = 1, 20
    net = fill_params(HSPose("PoseNet_only", chamfer_w=1.0)).to(cuda).train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    batch = {k: v.to(cuda) for k, v in synth_batch(8, 1028, seed=3, train=True).items()}
    torch.manual_seed(99)
    ctx = gcn3d.record_rf_indices(rf) if mode == "fp32" else gcn3d.force_rf_indices(rf)
    with ctx, torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        out, losses = net(**batch, do_loss=True)
    total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
    total.backward()
    res[mode] = ({n: p.grad.float().clone() for n, p in net.named_parameters() if p.grad is not None},
                 {k: v.item() for k, v in losses["fsnet_loss"].items()})
print(res["fp32"][1]); print(res["bf16"][1])
for n in res["fp32"][0]:
    a, b = res["fp32"][0][n].reshape(-1), res["bf16"][0][n].reshape(-1)
    print(f"{n:55s} |g32|={a.norm().item():.3e} |g16|={b.norm().item():.3e} cos={torch.nn.functional.cosine_similarity(a, b, dim=0).item():.4f}")
