"""On-device batched augmentation used by HSPose.data_augment.

Restates the math of the reference's datasets/data_augmentation.py
(`defor_3D_bb_in_batch` :70-79, `defor_3D_bc_in_batch` :106-127,
`defor_3D_pc` :134-140, `defor_3D_rt_in_batch` :183-190) as row-vector algebra
(points @ R instead of (R^T @ points^T)^T) and consumes the device RNG in the
same order and shapes as HSPose.data_augment (reference network/HSPose.py:233-246).
Out of the kernel scope: plain PyTorch.
"""
import torch


def _to_object_frame(pc, R, t):
    return (pc - t.unsqueeze(1)) @ R          # == (R^T (p - t))^T


def _to_camera_frame(pc_obj, R, t):
    return pc_obj @ R.transpose(1, 2) + t.unsqueeze(1)


def deform_bb(pc, model_point, R, t, s, sym, aug_bb):
    sym_aug = (aug_bb + aug_bb.flip(-1)) / 2.0   # == aug_bb[:, [2, 1, 0]] without a host index tensor
    scale = torch.where((sym[:, 0] == 1).unsqueeze(-1), sym_aug, aug_bb)
    pc_new = _to_camera_frame(_to_object_frame(pc, R, t) * scale.unsqueeze(1), R, t)
    return pc_new, s * scale, model_point * scale.unsqueeze(1)


def deform_rt(pc, R, t, aug_rt_t, aug_rt_r):
    pc_new = (pc + aug_rt_t.unsqueeze(1)) @ aug_rt_r.transpose(1, 2)
    t_new = ((t + aug_rt_t).unsqueeze(1) @ aug_rt_r.transpose(1, 2)).squeeze(1)
    return pc_new, aug_rt_r @ R, t_new


def deform_bc(pc, R, t, s, model_point, nocs_scale):
    bs = pc.size(0)
    ey_up = torch.rand((bs, 1), device=pc.device) * (1.2 - 0.8) + 0.8
    ey_down = torch.rand((bs, 1), device=pc.device) * (1.2 - 0.8) + 0.8
    s_y = s[..., 1].unsqueeze(-1)

    def taper(p):
        f = (p[..., 1] + s_y / 2.0) / s_y * (ey_up - ey_down) + ey_down
        return torch.stack([p[..., 0] * f, p[..., 1], p[..., 2] * f], dim=-1)

    pc_new = _to_camera_frame(taper(_to_object_frame(pc, R, t)), R, t)
    mp = taper(model_point)
    s_new = (mp.max(dim=1)[0] - mp.min(dim=1)[0]) * nocs_scale.unsqueeze(-1)
    return pc_new, s_new, ey_up, ey_down


def deform_pc(pc, gt_t, r):
    defor = torch.rand(pc.shape, device=pc.device) * r
    return pc + defor * (pc - gt_t.unsqueeze(1)), defor
