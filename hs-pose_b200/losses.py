"""Loss terms the stand-alone train step uses (BASELINE.json config 3:
"Chamfer + fs_net losses").

`fs_net_loss` restates the vectorised terms of the reference's
losses/fs_net_loss.py:31-76 (Rot1 :123, Rot2 :143, cosine :167/:191, regulariser
:222, Tran/Size :238-242, R_con :93-110) with the same FLAGS weights and the same
`valid_num` renormalisation for y-symmetric objects.  `chamfer_recon_loss` is the
sm_100a K7 kernel (the reference vendors an unused Chamfer extension,
tools/pyTorchChamferDistance/).  The other three reference loss groups
(recon_6face / geo / prop) are out of the kernel scope:
HSPose picks them up from the reference tree when it is importable.
"""
import torch
import torch.nn as nn

from . import ops
from .flags import FLAGS


def _dot(a, b):
    return (a * b).sum(dim=-1)


def get_gt_v(Rs):
    """Green / red axis targets (reference tools/training_utils.py:59-73, axis == 2):
    rows 1 and 2 of (R @ [[0,0,1],[0,1,0],[0,0,0]])^T, i.e. R e_y and R e_x."""
    return Rs[:, :, 1], Rs[:, :, 0]


class fs_net_loss(nn.Module):
    def __init__(self):
        super().__init__()
        if FLAGS.fsnet_loss_type == 'l1':
            mk = lambda beta: nn.L1Loss()
        elif FLAGS.fsnet_loss_type == 'smoothl1':
            mk = lambda beta: nn.SmoothL1Loss(beta=beta)
        else:
            raise NotImplementedError
        self.loss_func_t, self.loss_func_s = mk(0.5), mk(0.5)
        self.loss_func_Rot1, self.loss_func_Rot2 = mk(0.5), mk(0.5)
        self.loss_func_r_con, self.loss_func_Recon = mk(0.5), mk(0.3)

    @staticmethod
    def _renorm(res, flag, bs):
        valid = flag.sum()
        return torch.where(valid > 0, res * bs / valid.clamp(min=1), res)

    def forward(self, name_list, pred_list, gt_list, sym):
        out = {}
        nosym = sym[:, 0] == 0
        bs = sym.shape[0]
        if "Rot1" in name_list:
            out["Rot1"] = FLAGS.rot_1_w * self.loss_func_Rot1(pred_list["Rot1"], gt_list["Rot1"])
        if "Rot1_cos" in name_list:
            out["Rot1_cos"] = FLAGS.rot_1_w * ((1.0 - _dot(pred_list["Rot1"], gt_list["Rot1"])) * 2.0).mean()
        if "Rot2" in name_list:
            f = nosym.unsqueeze(-1)
            res = self.loss_func_Rot2(torch.where(f, pred_list["Rot2"], torch.zeros_like(pred_list["Rot2"])),
                                      torch.where(f, gt_list["Rot2"], torch.zeros_like(gt_list["Rot2"])))
            out["Rot2"] = FLAGS.rot_2_w * self._renorm(res, nosym, bs)
        if "Rot2_cos" in name_list:
            res = (1.0 - _dot(pred_list["Rot2"], gt_list["Rot2"])) * 2.0
            res = torch.where(nosym, res, torch.zeros_like(res)).mean()
            out["Rot2_cos"] = FLAGS.rot_2_w * self._renorm(res, nosym, bs)
        if "Rot_regular" in name_list:
            res = _dot(pred_list["Rot1"], pred_list["Rot2"]).abs()
            res = torch.where(nosym, res, torch.zeros_like(res)).mean()
            out["Rot_r_a"] = FLAGS.rot_regular * self._renorm(res, nosym, bs)
        if "Recon" in name_list:
            out["Recon"] = FLAGS.recon_w * self.loss_func_Recon(pred_list["Recon"], gt_list["Recon"])
        if "Tran" in name_list:
            out["Tran"] = FLAGS.tran_w * self.loss_func_t(pred_list["Tran"], gt_list["Tran"])
        if "Size" in name_list:
            out["Size"] = FLAGS.size_w * self.loss_func_s(pred_list["Size"], gt_list["Size"])
        if "R_con" in name_list:
            dg = torch.norm(pred_list["Rot1"] - gt_list["Rot1"], dim=-1)
            res_g = self.loss_func_r_con(torch.exp(-13.7 * dg * dg), pred_list["Rot1_f"])
            dr = torch.norm(pred_list["Rot2"] - gt_list["Rot2"], dim=-1)
            con_gt = torch.exp(-13.7 * dr * dr)
            zero = torch.zeros_like(con_gt)
            res_r = self.loss_func_r_con(torch.where(nosym, con_gt, zero),
                                         torch.where(nosym, pred_list["Rot2_f"], zero))
            out["R_con"] = FLAGS.r_con_w * (res_r + res_g)
        return out


def chamfer_recon_loss(recon, PC, weight=1.0):
    """Symmetric Chamfer distance between the reconstructed and the observed cloud (K7)."""
    d_a, d_b, _, _ = ops.chamfer(recon, PC)
    return weight * (d_a.mean() + d_b.mean())
