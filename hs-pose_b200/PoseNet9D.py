"""Mirror of the reference's network/fs_net_repo/PoseNet9D.py:14-52.

forward(points (bs,N,3), obj_id) -> (recon, face_normal, face_dis, face_f,
p_green_R, p_red_R, f_green_R, f_red_R, Pred_T, Pred_s), same order/shapes.
"""
import torch
import torch.nn as nn

from .FaceRecon import FaceRecon
from .PoseR import Rot_green, Rot_red
from .PoseTs import Pose_Ts
from . import ops
from .flags import FLAGS


class PoseNet9D(nn.Module):
    def __init__(self):
        super(PoseNet9D, self).__init__()
        self.rot_green = Rot_green()
        self.rot_red = Rot_red()
        self.face_recon = FaceRecon()
        self.ts = Pose_Ts()

    def forward(self, points, obj_id):
        bs, p_num = points.shape[0], points.shape[1]
        mean = points.mean(dim=1, keepdim=True)
        centred = points - mean
        # the three heads' first blocks read the same feature buffer as conv1d_block[0]: let the
        # backbone evaluate them as one autograd node on the mixed-precision path
        self.face_recon.joint_first = [(self.rot_green.conv1, self.rot_green.bn1),
                                       (self.rot_red.conv1, self.rot_red.bn1),
                                       (self.ts.conv1, self.ts.bn1)]
        recon, face, feat = self.face_recon(centred, obj_id)
        joint = self.face_recon.joint_out or [None, None, None]

        self.face_raw = None
        if FLAGS.train:
            recon, face = recon.float() + mean, face.float()
            self.face_raw = face      # (bs, N, 30) fp32: the fused loss kernel (K8) starts from the raw head output
            face_normal = face[:, :, :18].view(bs, p_num, 6, 3)
            face_normal = face_normal / torch.norm(face_normal, dim=-1, keepdim=True)
            face_dis = face[:, :, 18:24]
            face_f = torch.sigmoid(face[:, :, 24:])
        else:
            face_normal, face_dis, face_f, recon = [None] * 4
        # mixed precision: one bf16 buffer (bs, N, 1296) = [feat | centred xyz | 0] feeds all heads
        feat_pad = self.face_recon.feat_padded
        head_in = feat if feat_pad is None else feat_pad
        green_R_vec = self.rot_green.forward_points(head_in, joint[0]).float()   # b x 4
        red_R_vec = self.rot_red.forward_points(head_in, joint[1]).float()       # b x 4
        p_green_R = green_R_vec[:, 1:] / (torch.norm(green_R_vec[:, 1:], dim=1, keepdim=True) + 1e-6)
        p_red_R = red_R_vec[:, 1:] / (torch.norm(red_R_vec[:, 1:], dim=1, keepdim=True) + 1e-6)
        f_green_R = torch.sigmoid(green_R_vec[:, 0])
        f_red_R = torch.sigmoid(red_R_vec[:, 0])

        feat_for_ts = torch.cat([feat, centred], dim=2) if feat_pad is None else feat_pad
        T, s = self.ts.forward_points(feat_for_ts, joint[2])
        ops.flush_counters()      # BatchNorm num_batches_tracked of every fused block above: one multi-tensor add
        Pred_T = T.float() + mean[:, 0, :]
        Pred_s = s.float()
        return recon, face_normal, face_dis, face_f, p_green_R, p_red_R, f_green_R, f_red_R, Pred_T, Pred_s
