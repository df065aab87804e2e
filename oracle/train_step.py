"""CPU train step / forward of the reference's algorithm — TEST INFRASTRUCTURE.

Materialising PyTorch port (oracle/torch_oracle.py) of the path, driven as the
reference's engine/train.py:76-110 drives it: forward with batch-stat BatchNorm,
loss terms, backward, gradient clip (5), optimiser step.  Used ONLY by
bench.py's `cpu_baseline` leg and `--impl reference` arm (kind "port": the
reference itself is Python and its tree does not travel to the GPU box).
"""
import torch

from . import torch_oracle as to


def chamfer_materialised(a, b):
    d = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1)
    return d.min(dim=2)[0].mean() + d.min(dim=1)[0].mean()


class OracleTrainer:
    def __init__(self, state_dict, lr=1e-4, chamfer_w=1.0, k=20, S=7):
        # parameters (leaf tensors) and buffers, keyed as the reference state_dict
        self.sd = {}
        self.params = []
        for name, t in state_dict.items():
            t = t.detach().clone().float().cpu() if t.is_floating_point() else t.detach().clone().cpu()
            leaf = name.rsplit(".", 1)[-1]
            if t.is_floating_point() and leaf not in ("running_mean", "running_var"):
                t.requires_grad_(True)
                self.params.append(t)
            self.sd[name] = t
        self.opt = torch.optim.Adam(self.params, lr=lr)
        self.chamfer_w, self.k, self.S = chamfer_w, k, S

    def forward_eval(self, batch):
        with torch.no_grad():
            return to.posenet9d(self.sd, batch["PC"], batch["obj_id"], k=self.k, S=self.S, train=False)

    def step(self, batch):
        from hspose_b200.losses import fs_net_loss, get_gt_v   # pure torch, runs on CPU
        from hspose_b200.HSPose import control_loss
        out = to.posenet9d(self.sd, batch["PC"], batch["obj_id"], k=self.k, S=self.S, train=True,
                           bn_training=True, dropout_p=0.2)
        green, red = get_gt_v(batch["gt_R"])
        pred = {"Rot1": out["p_green_R"], "Rot1_f": out["f_green_R"], "Rot2": out["p_red_R"],
                "Rot2_f": out["f_red_R"], "Recon": out["recon"], "Tran": out["Pred_T"],
                "Size": out["Pred_s"]}
        gt = {"Rot1": green, "Rot2": red, "Recon": batch["PC"], "Tran": batch["gt_t"],
              "Size": batch["gt_s"]}
        losses = fs_net_loss()(control_loss("PoseNet_only")[0], pred, gt, batch["sym"])
        total = sum(v.reshape(()) for v in losses.values())
        if self.chamfer_w > 0:
            total = total + self.chamfer_w * chamfer_materialised(out["recon"], batch["PC"])
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 5.0)
        self.opt.step()
        return float(total)
