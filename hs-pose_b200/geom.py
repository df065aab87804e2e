"""Small per-object rotation algebra shared by the loss terms and the evaluation post-processing.

Restates (as closed-form vector algebra, no per-object Python loops, no host syncs, CUDA-graph safe):
  get_vertical_rot_vec_in_batch   reference tools/rot_utils.py:39-65
  get_rot_mat_y_first             reference tools/rot_utils.py:76-85
  to_R_matrices                   reference tools/rot_utils.py:95-98
  generate_RT                     reference tools/geom_utils.py:232-244
  get_gt_v                        reference tools/training_utils.py:59-73
"""
import math

import torch
import torch.nn.functional as F


def vertical_rot_vec(c1, c2, y, z):
    """Make the two predicted axes orthogonal, sharing the correction by confidence.

    c1, c2 (bs,) confidences; y, z (bs,3) unit vectors.  Both are rotated about r = y x z so that
    their angle becomes 90 degrees: y by theta_1 = c2/(c1+c2) (angle - pi/2), z by -theta_2 with
    theta_2 = c1/(c1+c2) (angle - pi/2).  Because r is orthogonal to y and z the Rodrigues matrix
    of the reference (rot_utils.py:67-74) reduces to  v' = cos(t) v + sin(t) (r x v)."""
    c1 = c1.unsqueeze(-1)
    c2 = c2.unsqueeze(-1)
    r = torch.cross(y, z, dim=-1)
    r = r / (torch.norm(r, dim=-1, keepdim=True) + 1e-8)
    cos_yz = torch.clamp(torch.sum(y * z, dim=-1, keepdim=True), -1 + 1e-6, 1 - 1e-6)
    excess = torch.acos(cos_yz) - math.pi / 2
    theta_1 = c2 / (c1 + c2) * excess
    theta_2 = c1 / (c1 + c2) * excess
    new_y = torch.cos(theta_1) * y + torch.sin(theta_1) * torch.cross(r, y, dim=-1)
    new_z = torch.cos(theta_2) * z - torch.sin(theta_2) * torch.cross(r, z, dim=-1)
    return new_y, new_z


def rot_mat_y_first(y, x):
    """(bs,3,3) rotation with columns (x', y', z'), y kept, x re-orthogonalised."""
    y = F.normalize(y, p=2, dim=-1)
    z = F.normalize(torch.cross(x, y, dim=-1), p=2, dim=-1)
    x = torch.cross(y, z, dim=-1)
    return torch.stack((x, y, z), dim=-1)


def to_R_matrices(f_g, f_r, p_g, p_r):
    new_y, new_x = vertical_rot_vec(f_g, f_r, p_g, p_r)
    return rot_mat_y_first(new_y, new_x)


def generate_RT(R, f, T, mode, sym):
    """(bs,4,4) pose matrices of the evaluation loop (evaluation/evaluate.py:106).

    mode == 'vec': R = [p_green (bs,3), p_red (bs,3)], f = [f_green (bs,), f_red (bs,)]; the red
    confidence of y-symmetric objects (sym[:,0] == 1) is zeroed.  mode == 'gt': R is (bs,3,3)."""
    bs = T.shape[0]
    res = torch.zeros(bs, 4, 4, dtype=T.dtype, device=T.device)
    if mode == "vec":
        f_green, f_red = f[0].reshape(-1), f[1].reshape(-1)
        f_red = torch.where(sym[:, 0] == 1, torch.zeros_like(f_red), f_red)
        Rs = to_R_matrices(f_green, f_red, R[0], R[1])
    else:
        Rs = R
    res[:, :3, :3] = Rs
    res[:, :3, 3] = T
    res[:, 3, 3] = 1.0
    return res


def get_gt_v(Rs):
    """Green / red axis targets (axis == 2): R e_y and R e_x."""
    return Rs[:, :, 1], Rs[:, :, 0]


def inv3x3(m):
    """Closed-form inverse of (...,3,3) matrices (adjugate / determinant): no library call, no
    host-side singularity check (torch.inverse synchronises), usable inside a CUDA graph."""
    a, b, c = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    d, e, f = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    g, h, i = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    A = e * i - f * h
    B = -(d * i - f * g)
    C = d * h - e * g
    det = a * A + b * B + c * C
    adj = torch.stack([
        torch.stack([A, -(b * i - c * h), b * f - c * e], dim=-1),
        torch.stack([B, a * i - c * g, -(a * f - c * d)], dim=-1),
        torch.stack([C, -(a * h - b * g), a * e - b * d], dim=-1)], dim=-2)
    return adj / det.unsqueeze(-1).unsqueeze(-1)
