"""Translation/size head — mirror of the reference's network/fs_net_repo/PoseTs.py:12-45."""
import torch.nn as nn

from .PoseR import _PointHead
from .flags import FLAGS


class Pose_Ts(_PointHead):
    def __init__(self):
        super().__init__(FLAGS.feat_c_ts, FLAGS.Ts_c)
        self.relu1 = nn.ReLU()
        self.relu2 = nn.ReLU()
        self.relu3 = nn.ReLU()

    def forward_points(self, x_bnc, first=None):
        x = super().forward_points(x_bnc, first)
        return x[:, 0:3], x[:, 3:6]
