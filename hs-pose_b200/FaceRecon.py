"""B200-native mirror of the reference's network/fs_net_repo/FaceRecon.py.

Same constructor-time FLAGS reads, sub-module names and state_dict keys
(FaceRecon.py:12-68) and the same forward contract (:70-128):
    forward(vertices (bs,N,3) centred, cat_id (bs,)|(bs,1)) -> (recon, face, feat)
The backbone runs on the fused sm_100a kernels of `gcn3d`; the geometric
neighbour tables are computed once per resolution; the three nearest-neighbour
up-samplings write straight into the (bs, N, 1286) concat buffer.
Dense 1x1-conv stacks are evaluated in (bs, N, C) layout as GEMMs — no
(bs, C, N) transposes.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gcn3d, ops
from .flags import FLAGS


def bn_points(bn: nn.BatchNorm1d, x_bnc):
    """nn.BatchNorm1d over the channel axis of a (bs, N, C) tensor — the same
    statistics as bn(x.transpose(1, 2)).transpose(1, 2) (FaceRecon.py:90-95)."""
    B, N, C = x_bnc.shape
    return bn(x_bnc.reshape(B * N, C)).view(B, N, C)


def conv1x1(conv: nn.Conv1d, x_bnc):
    """nn.Conv1d(kernel_size=1) applied in (bs, N, C) layout."""
    return F.linear(x_bnc, conv.weight[:, :, 0], conv.bias)


def seq_points(seq: nn.Sequential, x_bnc):
    """Run a Conv1d/BatchNorm1d/ReLU nn.Sequential (FaceRecon.py:38-68) in (bs, N, C) layout."""
    for m in seq:
        if isinstance(m, nn.Conv1d):
            x_bnc = conv1x1(m, x_bnc)
        elif isinstance(m, nn.BatchNorm1d):
            x_bnc = bn_points(m, x_bnc)
        elif isinstance(m, nn.ReLU):
            x_bnc = F.relu(x_bnc)
        else:
            raise NotImplementedError(type(m))
    return x_bnc


class FaceRecon(nn.Module):
    def __init__(self):
        super(FaceRecon, self).__init__()
        self.neighbor_num = FLAGS.gcn_n_num
        self.support_num = FLAGS.gcn_sup_num

        self.conv_0 = gcn3d.HSlayer_surface(kernel_num=128, support_num=self.support_num)
        self.conv_1 = gcn3d.HS_layer(128, 128, support_num=self.support_num)
        self.pool_1 = gcn3d.Pool_layer(pooling_rate=4, neighbor_num=4)
        self.conv_2 = gcn3d.HS_layer(128, 256, support_num=self.support_num)
        self.conv_3 = gcn3d.HS_layer(256, 256, support_num=self.support_num)
        self.pool_2 = gcn3d.Pool_layer(pooling_rate=4, neighbor_num=4)
        self.conv_4 = gcn3d.HS_layer(256, 512, support_num=self.support_num)

        self.bn1 = nn.BatchNorm1d(128)
        self.bn2 = nn.BatchNorm1d(256)
        self.bn3 = nn.BatchNorm1d(256)

        self.recon_num = 3
        self.face_recon_num = FLAGS.face_recon_c
        self.obj_c = FLAGS.obj_c
        dim_fuse = sum([128, 128, 256, 256, 512, FLAGS.obj_c])

        if FLAGS.train:
            self.conv1d_block = nn.Sequential(
                nn.Conv1d(dim_fuse, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                nn.Conv1d(512, 256, 1), nn.BatchNorm1d(256), nn.ReLU(inplace=True),
            )
            self.recon_head = nn.Sequential(
                nn.Conv1d(256, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                nn.Conv1d(128, self.recon_num, 1),
            )
            self.face_head = nn.Sequential(
                nn.Conv1d(FLAGS.feat_face + 3, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                nn.Conv1d(512, 256, 1), nn.BatchNorm1d(256), nn.ReLU(inplace=True),
                nn.Conv1d(256, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                nn.Conv1d(128, self.face_recon_num, 1),
            )

    def forward(self, vertices: "tensor (bs, vetice_num, 3)", cat_id: "tensor (bs, 1)"):
        bs, vertice_num, _ = vertices.size()
        if cat_id.shape[0] == 1:
            obj_idh = cat_id.view(-1, 1).repeat(cat_id.shape[0], 1)
        else:
            obj_idh = cat_id.view(-1, 1)
        one_hot = torch.zeros(bs, self.obj_c, device=vertices.device).scatter_(
            1, obj_idh.to(vertices.device).long(), 1)

        k = self.neighbor_num
        vertices = vertices.contiguous()
        with gcn3d.neighbor_cache():
            fm_0 = F.relu(self.conv_0(vertices, k))
            fm_1 = F.relu(bn_points(self.bn1, self.conv_1(vertices, fm_0, k)))
            v_pool_1, fm_pool_1 = self.pool_1(vertices, fm_1)
            k1 = min(k, v_pool_1.shape[1] // 8)
            fm_2 = F.relu(bn_points(self.bn2, self.conv_2(v_pool_1, fm_pool_1, k1)))
            fm_3 = F.relu(bn_points(self.bn3, self.conv_3(v_pool_1, fm_2, k1)))
            v_pool_2, fm_pool_2 = self.pool_2(v_pool_1, fm_3)
            fm_4 = self.conv_4(v_pool_2, fm_pool_2, min(k, v_pool_2.shape[1] // 8))
        f_global = fm_4.max(1)[0]  # (bs, 512)

        # nearest up-sampling (FaceRecon.py:100-104) fused with the concat (:107)
        nearest_pool_1 = ops.knn3(vertices, v_pool_1, 1, drop_first=0, formula=ops.DIST_NEAREST)[1]
        nearest_pool_2 = ops.knn3(vertices, v_pool_2, 1, drop_first=0, formula=ops.DIST_NEAREST)[1]
        feat = ops.concat_upsample(
            [fm_0, fm_1, fm_2, fm_3, fm_4, one_hot],
            [None, None, nearest_pool_1[..., 0], nearest_pool_1[..., 0], nearest_pool_2[..., 0], "bcast"],
            vertice_num)

        if FLAGS.train:
            conv1d_out = seq_points(self.conv1d_block, feat)           # (bs, N, 256)
            recon = seq_points(self.recon_head, conv1d_out)            # (bs, N, 3)
            feat_face_in = torch.cat(
                [f_global.unsqueeze(1).expand(-1, vertice_num, -1), conv1d_out, vertices], dim=2)
            face = seq_points(self.face_head, feat_face_in)            # (bs, N, 30)
            return recon, face, feat
        return None, None, feat
