#!/usr/bin/env python
"""Debug: per-parameter gradient diff of the teacher-forced train step vs the golden."""
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
g = dict(np.load(os.path.join(ROOT, "tests/golden/e2e_train.npz")))
F = hf.get_flags()
for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro"):
    setattr(F, n, 0.0)
F.train, F.gcn_n_num = 1, 20
net = fill_params(HSPose("PoseNet_only")).to(cuda).train()
for m in net.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
batch = {k: v.to(cuda) for k, v in synth_batch(4, 1028, seed=2, train=True).items()}
rf = [torch.from_numpy(g[f"rf{i}"].astype(np.int64)) for i in range(4)]
torch.manual_seed(4321)
for shape in [(4, 1)] * 6 + [(4, 1028, 3)]:
    torch.rand(shape)
with gcn3d.force_rf_indices(rf):
    out, losses = net(**batch, do_loss=True)
total = sum(v.reshape(()) for v in losses["fsnet_loss"].values())
total.backward()
params = dict(net.named_parameters())
for key in sorted(g):
    if key.startswith("grad::"):
        p = params[key[6:]].grad.cpu().numpy()
        ref = g[key]
        d = np.abs(p - ref)
        tol = 2e-4 * max(1.0, float(np.abs(ref).max()))
        print(f"{key[6:]:60s} max|ref|={np.abs(ref).max():.3e} maxdiff={d.max():.3e} bad={(d > tol).mean():.4f} "
              f"cos={float((p*ref).sum()/ (np.linalg.norm(p)*np.linalg.norm(ref)+1e-30)):.6f}")

# ---- oracle on the GPU, fp32 and fp64, same forced RF tables and pool samples
from oracle import torch_oracle as to
from hspose_b200.losses import fs_net_loss, get_gt_v
from hspose_b200.HSPose import control_loss
torch.manual_seed(4321)
for shape in [(4, 1)] * 6 + [(4, 1028, 3)]:
    torch.rand(shape)
s1 = torch.randperm(1028)[:257]; s2 = torch.randperm(257)[:64]

def oracle_grads(dtype):
    sd = {}
    for name, t in net.state_dict().items():
        t = t.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype)
            if name.rsplit(".", 1)[-1] not in ("running_mean", "running_var"):
                t.requires_grad_(True)
        sd[name] = t
    b = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}
    out = to.posenet9d(sd, b["PC"], b["obj_id"], k=20, S=7, train=True, bn_training=True,
                       samples=(s1.to(cuda), s2.to(cuda)), rf_indices=[r.to(cuda) for r in rf])
    green, red = get_gt_v(b["gt_R"])
    pred = {"Rot1": out["p_green_R"], "Rot1_f": out["f_green_R"], "Rot2": out["p_red_R"],
            "Rot2_f": out["f_red_R"], "Recon": out["recon"], "Tran": out["Pred_T"], "Size": out["Pred_s"]}
    gt = {"Rot1": green, "Rot2": red, "Recon": b["PC"], "Tran": b["gt_t"], "Size": b["gt_s"]}
    losses = fs_net_loss()(control_loss("PoseNet_only")[0], pred, gt, b["sym"])
    tot = sum(v.reshape(()) for v in losses.values())
    tot.backward()
    return {n: t.grad.double().cpu().numpy() for n, t in sd.items() if t.is_floating_point() and t.grad is not None}

g32 = oracle_grads(torch.float32)
g64 = oracle_grads(torch.float64)
def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
print(f"{'param':50s} ours-o64  gold-o64  o32-o64   ours-gold")
for key in sorted(g):
    if key.startswith("grad::") and "face_recon.conv" in key or key.startswith("grad::posenet.face_recon.bn"):
        n = key[6:]
        mine = params[n].grad.double().cpu().numpy()
        print(f"{n[8:]:50s} {rel(mine, g64[n]):.2e}  {rel(g[key], g64[n]):.2e}  {rel(g32[n], g64[n]):.2e}  {rel(mine, g[key]):.2e}")
