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


FEAT_LD = 1296   # 1286 feature channels + 3 centred xyz + 7 zeros: 16-element aligned rows


def mixed_precision():
    """The fast dense path is active inside torch.autocast('cuda', bf16): bf16 activations in
    16-byte aligned (M,C) matrices for the tensor-core GEMMs, fused BN+ReLU kernels (K6b)."""
    return torch.is_autocast_enabled("cuda")


def bn_points(bn: nn.BatchNorm1d, x_bnc, relu=False):
    """nn.BatchNorm1d over the channel axis of a (bs, N, C) / (M, C) tensor — the same
    statistics as bn(x.transpose(1, 2)).transpose(1, 2) (FaceRecon.py:90-95) — optionally
    fused with the ReLU that follows it everywhere in the reference."""
    shape = x_bnc.shape
    x2 = x_bnc.reshape(-1, shape[-1])
    if bn.training and mixed_precision() and x2.is_cuda and shape[-1] % 8 == 0 and x2.shape[0] > 1:
        ops.bump_counter(bn.num_batches_tracked)
        y = ops.bn_relu(x2, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                        bn.momentum, relu)
    else:
        y = bn(x2)
        if relu:
            y = F.relu(y)
    return y.view(shape)


def conv1x1(conv: nn.Conv1d, x_bnc, pad_in=0, pad_out=0):
    """nn.Conv1d(kernel_size=1) applied in (bs, N, C) layout.  pad_in / pad_out append
    zero input columns / output rows to the weight so GEMM dims stay 16-byte aligned."""
    w, b = conv.weight.squeeze(-1), conv.bias
    if mixed_precision() and x_bnc.is_cuda:
        # K6 tensor-core GEMM (forward, dgrad, wgrad); the kernel pads / clips ragged widths itself
        if pad_in:
            w = F.pad(w, (0, pad_in))
        y = ops.linear_tc(x_bnc, w, b)
        return F.pad(y, (0, pad_out)) if pad_out else y
    if pad_in or pad_out:
        w = F.pad(w, (0, pad_in, 0, pad_out))
        if b is not None and pad_out:
            b = F.pad(b, (0, pad_out))
    return F.linear(x_bnc, w, b)


def _eval_folded(conv: nn.Conv1d, bn: nn.BatchNorm1d, x_bnc, relu):
    """Evaluation (running statistics), fp32: Conv1d(k=1) -> BatchNorm1d -> ReLU folded into ONE fp32-accurate
    tensor-core GEMM (ops.linear_fp32x: 3-way bf16 split, K6 kernel, bias + ReLU in the epilogue).  The folded
    and pre-split weights are cached on the conv module and rebuilt when any source tensor changes."""
    key = tuple((t.data_ptr(), t._version) for t in (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                                     bn.running_var) if t is not None) + (x_bnc.shape[-1],)
    cache = getattr(conv, "_hsp_eval_fold", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = conv.weight.squeeze(-1) * s[:, None]
            pad_in = x_bnc.shape[-1] - conv.in_channels
            if pad_in:
                w = F.pad(w, (0, pad_in))
            b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
            b = ((b0 - bn.running_mean) * s + bn.bias).float().contiguous()
            kpad = (w.shape[1] + 63) // 64 * 64
            cache = (key, ops.split_bf16(w.float().contiguous(), ops._SPLIT_B, kpad), b)
        conv._hsp_eval_fold = cache
    shape = x_bnc.shape
    y = ops.eval_linear(x_bnc.reshape(-1, shape[-1]), cache[1], cache[2], relu)
    return y.view(*shape[:-1], conv.out_channels)


def conv_bn_relu_points(conv: nn.Conv1d, bn: nn.BatchNorm1d, x_bnc, relu=True):
    """Conv1d(k=1) -> BatchNorm1d -> ReLU on a (bs, N, C) / (M, C) tensor.  On the mixed-precision
    training path this is ONE fused autograd node (ops.linear_bn_relu: bf16 tensor-core GEMMs,
    K6b BN kernels, bias gradient from the BN backward); otherwise the three modules in turn."""
    shape = x_bnc.shape
    pad_in = shape[-1] - conv.in_channels
    if (bn.training and mixed_precision() and x_bnc.is_cuda and conv.out_channels % 8 == 0
            and shape[-1] % 8 == 0 and x_bnc.numel() // shape[-1] > 1):
        w = conv.weight.squeeze(-1)
        if pad_in:
            w = F.pad(w, (0, pad_in))
        ops.bump_counter(bn.num_batches_tracked)
        z = ops.linear_bn_relu(x_bnc.reshape(-1, shape[-1]), w, conv.bias, bn.weight, bn.bias,
                               bn.running_mean, bn.running_var, bn.eps, bn.momentum, relu)
        return z.view(*shape[:-1], conv.out_channels)
    if (not bn.training and not mixed_precision() and x_bnc.is_cuda and x_bnc.dtype == torch.float32
            and conv.in_channels >= 512 and conv.out_channels % 4 == 0 and x_bnc.numel() // shape[-1] >= 1024
            and bn.running_mean is not None and bn.affine):
        return _eval_folded(conv, bn, x_bnc, relu)
    return bn_points(bn, conv1x1(conv, x_bnc, pad_in), relu=relu)


def fused_block_ok(conv: nn.Conv1d, bn: nn.BatchNorm1d, x_bnc):
    shape = x_bnc.shape
    return (bn.training and mixed_precision() and x_bnc.is_cuda and conv.out_channels % 8 == 0
            and shape[-1] % 8 == 0 and x_bnc.numel() // shape[-1] > 1)


def multi_conv_bn_relu_points(pairs, x_bnc):
    """[(conv, bn)] all reading x_bnc -> list of outputs; one autograd node on the mixed path
    (ops.multi_linear_bn_relu), else the blocks one by one."""
    shape = x_bnc.shape
    if not all(fused_block_ok(c, b, x_bnc) for c, b in pairs):
        return [conv_bn_relu_points(c, b, x_bnc) for c, b in pairs]
    blocks = []
    for conv, bn in pairs:
        w = conv.weight.squeeze(-1)
        pad_in = shape[-1] - conv.in_channels
        if pad_in:
            w = F.pad(w, (0, pad_in))
        ops.bump_counter(bn.num_batches_tracked)
        blocks.append((w, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                       bn.momentum, True))
    outs = ops.multi_linear_bn_relu(x_bnc.reshape(-1, shape[-1]), blocks)
    return [z.view(*shape[:-1], c.out_channels) for z, (c, _) in zip(outs, pairs)]


def seq_points(seq: nn.Sequential, x_bnc):
    """Run a Conv1d/BatchNorm1d/ReLU nn.Sequential (FaceRecon.py:38-68) in (bs, N, C) layout;
    BN+ReLU pairs are fused; a narrow last conv (3 / 30 channels) is computed 8-aligned."""
    mods = list(seq)   # an nn.Sequential or a plain list of its modules
    i = 0
    while i < len(mods):
        m = mods[i]
        if (isinstance(m, nn.Conv1d) and i + 2 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d)
                and isinstance(mods[i + 2], nn.ReLU)):
            x_bnc = conv_bn_relu_points(m, mods[i + 1], x_bnc, relu=True)
            i += 3
            continue
        if isinstance(m, nn.Conv1d):
            pad_in = x_bnc.shape[-1] - m.in_channels
            x_bnc = conv1x1(m, x_bnc, pad_in)
        elif isinstance(m, nn.BatchNorm1d):
            fuse = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            x_bnc = bn_points(m, x_bnc, relu=fuse)
            i += 1 if fuse else 0
        elif isinstance(m, nn.ReLU):
            x_bnc = F.relu(x_bnc)
        else:
            raise NotImplementedError(type(m))
        i += 1
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
            fm_1 = bn_points(self.bn1, self.conv_1(vertices, fm_0, k), relu=True)
            v_pool_1, fm_pool_1 = self.pool_1(vertices, fm_1)
            k1 = min(k, v_pool_1.shape[1] // 8)
            fm_2 = bn_points(self.bn2, self.conv_2(v_pool_1, fm_pool_1, k1), relu=True)
            fm_3 = bn_points(self.bn3, self.conv_3(v_pool_1, fm_2, k1), relu=True)
            v_pool_2, fm_pool_2 = self.pool_2(v_pool_1, fm_3)
            fm_4 = self.conv_4(v_pool_2, fm_pool_2, min(k, v_pool_2.shape[1] // 8))
        f_global = fm_4.max(1)[0]  # (bs, 512)

        # nearest up-sampling (FaceRecon.py:100-104) fused with the concat (:107)
        nearest_pool_1 = ops.knn3(vertices, v_pool_1, 1, drop_first=0, formula=ops.DIST_NEAREST)[1]
        nearest_pool_2 = ops.knn3(vertices, v_pool_2, 1, drop_first=0, formula=ops.DIST_NEAREST)[1]
        pieces = [fm_0, fm_1, fm_2, fm_3, fm_4, one_hot]
        nns = [None, None, nearest_pool_1[..., 0], nearest_pool_1[..., 0], nearest_pool_2[..., 0], "bcast"]
        mixed = mixed_precision()
        if mixed:
            # bf16, rows padded to FEAT_LD; columns 1286:1289 carry the centred xyz so the same
            # buffer is PoseNet9D's `feat_for_ts` (PoseNet9D.py:47) — no second concat
            feat_pad = ops.concat_upsample(pieces + [vertices], nns + [None], vertice_num,
                                           ld=FEAT_LD, out_dtype=torch.bfloat16)
            feat = feat_pad[:, :, :feat_pad.shape[2] - (FEAT_LD - 1286)]
            self.feat_padded = feat_pad
        else:
            feat = ops.concat_upsample(pieces, nns, vertice_num)
            self.feat_padded = None

        # heads that read the same feature buffer (PoseNet9D registers them in `joint_first`)
        # share one autograd node with conv1d_block[0] on the mixed-precision path
        joint = list(getattr(self, "joint_first", None) or []) if mixed else []
        self.joint_out = None
        if FLAGS.train:
            if joint and fused_block_ok(self.conv1d_block[0], self.conv1d_block[1], feat_pad):
                outs = multi_conv_bn_relu_points(joint + [(self.conv1d_block[0], self.conv1d_block[1])], feat_pad)
                self.joint_out = outs[:-1]
                conv1d_out = seq_points(list(self.conv1d_block)[3:], outs[-1])             # (bs, N, 256)
            else:
                conv1d_out = seq_points(self.conv1d_block, feat_pad if mixed else feat)   # (bs, N, 256)
            recon = seq_points(self.recon_head, conv1d_out)                            # (bs, N, 3)
            face = self._face_head(f_global, conv1d_out, vertices)                     # (bs, N, 30)
            ops.flush_counters()
            return recon, face, feat
        if joint and all(fused_block_ok(c, b, feat_pad) for c, b in joint):
            self.joint_out = multi_conv_bn_relu_points(joint, feat_pad)
        ops.flush_counters()
        return None, None, feat

    def _face_head(self, f_global, conv1d_out, vertices):
        """face_head on cat[f_global repeated, conv1d_out, xyz] (FaceRecon.py:118-124).  The
        f_global block of the first conv is a per-object constant: W[:, :512] @ f_global[b]
        is computed once per object and broadcast instead of multiplying N identical rows."""
        bs, n, _ = conv1d_out.shape
        if not mixed_precision():
            x = torch.cat([f_global.unsqueeze(1).expand(-1, n, -1), conv1d_out, vertices], dim=2)
            return seq_points(self.face_head, x)
        conv0 = self.face_head[0]
        w = conv0.weight.squeeze(-1)
        cg = f_global.shape[1]
        per_obj = ops.linear_tc(f_global, w[:, :cg], conv0.bias).float()         # (bs, 512)
        pad = (-(conv1d_out.shape[2] + 3)) % 8
        tail = torch.cat([conv1d_out, vertices.to(conv1d_out.dtype),
                          conv1d_out.new_zeros(bs, n, pad)], dim=2)              # (bs, N, 264): one pass, 16-B rows
        bn0 = self.face_head[1]
        if bn0.training and n > 1:
            # face_head[0..2] = conv -> BN -> ReLU as one K6/K6b node; the f_global block enters the GEMM
            # epilogue as a per-object bias (no (bs, N, 512) broadcast add)
            ops.bump_counter(bn0.num_batches_tracked)
            z = ops.linear_bn_relu(tail.view(bs * n, -1), F.pad(w[:, cg:], (0, pad)), None, bn0.weight, bn0.bias,
                                   bn0.running_mean, bn0.running_var, bn0.eps, bn0.momentum, True,
                                   bias_rows=per_obj.contiguous(), rows_per_group=n)
            return seq_points(list(self.face_head)[3:], z.view(bs, n, -1))
        x = ops.linear_tc(tail, F.pad(w[:, cg:], (0, pad))) + per_obj.unsqueeze(1)
        return seq_points(list(self.face_head)[1:], x)
