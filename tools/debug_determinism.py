"""Run the same train forward/backward several times and report which gradients are not reproducible."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hspose_b200 import parallel
from hspose_b200.HSPose import HSPose
from hspose_b200.flags import get_flags
from hspose_b200.synth import fill_params, synth_batch

amp = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda")
F = get_flags()
for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro"):
    setattr(F, n, 0.0)
model = fill_params(HSPose("PoseNet_only", chamfer_w=1.0)).to(dev).train()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
batch = {k: v.to(dev) for k, v in synth_batch(B, 1028, seed=9, train=True).items()}
names = [n for n, p in model.posenet.named_parameters()]
runs = []
outs = []
for r in range(3):
    parallel.seed_all(4321)
    for p in model.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        out, losses = model(**batch, do_loss=True)
    total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
    total.backward()
    runs.append({n: p.grad.detach().clone() for n, p in model.posenet.named_parameters() if p.grad is not None})
    outs.append((total.item(), out["recon"].detach().clone(), out["p_green_R"].detach().clone()))
print("loss", [o[0] for o in outs], "fwd recon bitwise equal:", torch.equal(outs[0][1], outs[1][1]),
      torch.equal(outs[0][2], outs[1][2]))
rows = []
for n in runs[0]:
    a, b = runs[0][n], runs[1][n]
    rows.append(((a - b).norm().item() / (a.norm().item() + 1e-30), n, a.norm().item()))
rows.sort(reverse=True)
for r in rows[:25]:
    print(f"{r[0]:.3e}  {r[1]:55s} |g|={r[2]:.3e}")
ta = torch.cat([v.reshape(-1) for v in runs[0].values()]); tb = torch.cat([v.reshape(-1) for v in runs[1].values()])
print("total rel l2", ((ta - tb).norm() / ta.norm()).item())
