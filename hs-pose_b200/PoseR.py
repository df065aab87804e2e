"""Rotation heads — mirror of the reference's network/fs_net_repo/PoseR.py:10-70.

Same sub-module names / shapes (conv1..4 nn.Conv1d k=1, bn1..3, drop1) so the
state_dict is interchangeable.  forward() takes the reference's (bs, C, N)
layout; forward_points() takes (bs, N, C) and is what PoseNet9D calls — the
1x1 convolutions are GEMMs over points, no transposes.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .FaceRecon import conv_bn_relu_points, mixed_precision
from .flags import FLAGS


class _PointHead(nn.Module):
    def __init__(self, f, k):
        super().__init__()
        self.f = f
        self.k = k
        self.conv1 = torch.nn.Conv1d(self.f, 1024, 1)
        self.conv2 = torch.nn.Conv1d(1024, 256, 1)
        self.conv3 = torch.nn.Conv1d(256, 256, 1)
        self.conv4 = torch.nn.Conv1d(256, self.k, 1)
        self.drop1 = nn.Dropout(0.2)
        self.bn1 = nn.BatchNorm1d(1024)
        self.bn2 = nn.BatchNorm1d(256)
        self.bn3 = nn.BatchNorm1d(256)

    def forward_points(self, x_bnc, first=None):
        # x_bnc may be the 16-aligned padded feature buffer (extra columns get zero weights);
        # `first` is relu(bn1(conv1(x))) when PoseNet9D already computed it jointly with the others
        x = first if first is not None else conv_bn_relu_points(self.conv1, self.bn1, x_bnc)
        x = conv_bn_relu_points(self.conv2, self.bn2, x)
        if x.is_cuda and x.shape[-1] % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16):
            x = ops.colmax(x)                                   # (bs, 256): max over points
        else:
            x = torch.max(x, 1)[0]
        lin = ops.linear_tc if (mixed_precision() and x.is_cuda) else F.linear   # K6 on the mixed-precision path
        x = F.relu(self.bn3(lin(x, self.conv3.weight.squeeze(-1), self.conv3.bias)))
        x = self.drop1(x)
        x = lin(x, self.conv4.weight.squeeze(-1), self.conv4.bias)
        ops.flush_counters()
        return x.contiguous()

    def forward(self, x):
        """x: (bs, C, N) as in the reference (PoseR.py:26-39)."""
        return self.forward_points(x.transpose(1, 2))


class Rot_green(_PointHead):
    def __init__(self):
        super().__init__(FLAGS.feat_c_R, FLAGS.R_c)


class Rot_red(_PointHead):
    def __init__(self):
        super().__init__(FLAGS.feat_c_R, FLAGS.R_c)
