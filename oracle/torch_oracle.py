"""Materialising PyTorch restatement of the HS-Pose hot path — TEST INFRASTRUCTURE.

A functional (state_dict-driven) re-derivation of what the reference computes,
written with plain dense torch ops exactly where the reference materialises
tensors, so that (a) run on CPU it reproduces the reference's numbers (pinned
by tests/golden/*.npz, generated from the real reference), (b) autograd through
it is the gradient oracle for the CUDA backward kernels, and (c) timed on the
host cores it is bench.py's `cpu_baseline` ("port").  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may import it.

Reference lines each function follows are cited as gcn3d.py:L (=
network/fs_net_repo/gcn3d.py), FaceRecon.py:L, PoseNet9D.py:L, PoseR.py:L,
PoseTs.py:L.
"""
import torch
import torch.nn.functional as F

OBJ_C = 6  # FLAGS.obj_c


# ------------------------------------------------------------------ indices
def pairwise_neighbor_dist(v):
    """gcn3d.py:19-21 — ((-2*inner) + q_j) + q_i, fully materialised (B,N,N)."""
    inner = torch.bmm(v, v.transpose(1, 2))
    q = (v * v).sum(dim=2)
    return inner * (-2) + q[:, None, :] + q[:, :, None]


def neighbor_index(v, k):
    """gcn3d.py:15-24."""
    d = pairwise_neighbor_dist(v)
    return d.topk(k + 1, dim=-1, largest=False)[1][..., 1:]


def pairwise_nearest_dist(target, source):
    """gcn3d.py:31-34 — (s_j + t_i) - 2*inner."""
    inner = torch.bmm(target, source.transpose(1, 2))
    s2 = (source * source).sum(dim=2)
    t2 = (target * target).sum(dim=2)
    return s2[:, None, :] + t2[:, :, None] - 2 * inner


def nearest_index(target, source):
    """gcn3d.py:27-36 -> (B,N1,1)."""
    return pairwise_nearest_dist(target, source).topk(1, dim=-1, largest=False)[1]


def take_rows(t, idx):
    """gcn3d.py:39-47 — t (B,N,C), idx (B,M,n) -> (B,M,n,C)."""
    B, M, n = idx.shape
    flat = idx.reshape(B, M * n, 1).expand(-1, -1, t.shape[2])
    return torch.gather(t, 1, flat).view(B, M, n, t.shape[2])


def direction_norm(xyz, idx):
    """gcn3d.py:49-59."""
    return F.normalize(take_rows(xyz, idx) - xyz[:, :, None, :], dim=-1)


# ------------------------------------------------------------------ fused ops, materialised
def surface_graph_conv(xyz, idx, directions, S, C):
    """gcn3d.py:92-107."""
    B, N, k = idx.shape
    theta = torch.relu(direction_norm(xyz, idx) @ F.normalize(directions, dim=0))
    return theta.view(B, N, k, S, C).amax(dim=2).mean(dim=2)


def hs_graph_conv(xyz, idx, fm, weights, bias, directions, S, C):
    """gcn3d.py:158-181."""
    B, N, k = idx.shape
    theta = torch.relu(direction_norm(xyz, idx) @ F.normalize(directions, dim=0))
    P = fm @ weights + bias
    centre, support = P[..., :C], P[..., C:]
    act = (theta * take_rows(support, idx)).view(B, N, k, S, C)
    return centre + act.amax(dim=2).mean(dim=2)


def orl_global(feat, xyz, k):
    """gcn3d.py:211-218 -> (B,C) (the reference repeats it over N)."""
    return take_rows(feat, neighbor_index(xyz, k)).amax(dim=2).mean(dim=1)


def orl_fuse(feat, xyz, k, conv2_w):
    """ORL_forward, gcn3d.py:109-113 / :183-187 (conv2 has no bias)."""
    G = orl_global(feat, xyz, k)[:, None, :].expand(-1, feat.shape[1], -1)
    return torch.cat([feat, G], dim=-1) @ conv2_w[:, :, 0].t() + feat


# ------------------------------------------------------------------ layers
def surface_layer(sd, pre, xyz, k, S):
    """HSlayer_surface.forward, gcn3d.py:79-90."""
    C = sd[pre + "STE_layer.weight"].shape[0]
    ste = xyz @ sd[pre + "STE_layer.weight"][:, :, 0].t()
    g = surface_graph_conv(xyz, neighbor_index(xyz, k), sd[pre + "directions"], S, C)
    return orl_fuse(g, xyz, k, sd[pre + "conv2.weight"]) + ste


def hs_layer(sd, pre, xyz, fm, k, S, rf_idx=None):
    """HS_layer.forward, gcn3d.py:143-156.  rf_idx teacher-forces the RF-F
    neighbour table (protocol T2)."""
    C = sd[pre + "STE_layer.weight"].shape[0]
    ste = fm @ sd[pre + "STE_layer.weight"][:, :, 0].t()
    if rf_idx is None:
        rf_idx = neighbor_index(fm, k)
    g = hs_graph_conv(xyz, rf_idx, fm, sd[pre + "weights"], sd[pre + "bias"],
                      sd[pre + "directions"], S, C)
    return orl_fuse(g, xyz, k, sd[pre + "conv2.weight"]) + ste, rf_idx


def pool_layer(xyz, fm, sample_idx, rate_k=4):
    """Pool_layer.forward, gcn3d.py:226-246 with the randperm sample injected."""
    pooled = take_rows(fm, neighbor_index(xyz, rate_k)).amax(dim=2)
    return xyz[:, sample_idx], pooled[:, sample_idx]


def _bn(sd, pre, x_bnc, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm1d over (B,N,C) applied on the channel axis."""
    y = F.batch_norm(x_bnc.transpose(1, 2), sd[pre + "running_mean"], sd[pre + "running_var"],
                     sd[pre + "weight"], sd[pre + "bias"], training, momentum, eps)
    return y.transpose(1, 2)


def _conv1(sd, pre, x_bnc):
    """nn.Conv1d(kernel_size=1) on (B,N,C) layout."""
    y = x_bnc @ sd[pre + "weight"][:, :, 0].t()
    if pre + "bias" in sd:
        y = y + sd[pre + "bias"]
    return y


def face_recon(sd, xyz, cat_id, k=20, S=7, train=True, bn_training=False, samples=None,
               rf_indices=None, pre="posenet.face_recon."):
    """FaceRecon.forward, FaceRecon.py:70-128.
    samples: the two randperm prefixes (len N//4 and N//16) — the reference draws
    them from the CPU generator (gcn3d.py:243).  Returns dict with feat, recon,
    face, the RF-F index tensors and intermediate feature maps."""
    B, N, _ = xyz.shape
    one_hot = torch.zeros(B, OBJ_C, dtype=xyz.dtype, device=xyz.device)
    one_hot.scatter_(1, cat_id.view(-1, 1).long(), 1)
    rf = list(rf_indices) if rf_indices is not None else [None] * 4
    out_rf = []
    fm0 = torch.relu(surface_layer(sd, pre + "conv_0.", xyz, k, S))
    y, r = hs_layer(sd, pre + "conv_1.", xyz, fm0, k, S, rf[0]); out_rf.append(r)
    fm1 = torch.relu(_bn(sd, pre + "bn1.", y, bn_training))
    if samples is None:
        samples = (torch.randperm(N)[: int(N / 4)], None)
    v1, fp1 = pool_layer(xyz, fm1, samples[0])
    N1 = v1.shape[1]
    k1 = min(k, N1 // 8)
    y, r = hs_layer(sd, pre + "conv_2.", v1, fp1, k1, S, rf[1]); out_rf.append(r)
    fm2 = torch.relu(_bn(sd, pre + "bn2.", y, bn_training))
    y, r = hs_layer(sd, pre + "conv_3.", v1, fm2, k1, S, rf[2]); out_rf.append(r)
    fm3 = torch.relu(_bn(sd, pre + "bn3.", y, bn_training))
    s2 = samples[1] if samples[1] is not None else torch.randperm(N1)[: int(N1 / 4)]
    v2, fp2 = pool_layer(v1, fm3, s2)
    k2 = min(k, v2.shape[1] // 8)
    fm4, r = hs_layer(sd, pre + "conv_4.", v2, fp2, k2, S, rf[3]); out_rf.append(r)
    f_global = fm4.amax(dim=1)
    nn1 = nearest_index(xyz, v1)
    nn2 = nearest_index(xyz, v2)
    up2 = take_rows(fm2, nn1).squeeze(2)
    up3 = take_rows(fm3, nn1).squeeze(2)
    up4 = take_rows(fm4, nn2).squeeze(2)
    feat = torch.cat([fm0, fm1, up2, up3, up4, one_hot[:, None, :].expand(-1, N, -1)], dim=2)
    res = {"feat": feat, "rf_indices": out_rf, "fm": (fm0, fm1, fm2, fm3, fm4),
           "recon": None, "face": None, "samples": (samples[0], s2)}
    if train:
        x = feat
        for i in (0, 3, 6):  # conv1d_block: Conv-BN-ReLU x3 (FaceRecon.py:38-48)
            x = torch.relu(_bn(sd, f"{pre}conv1d_block.{i + 1}.",
                               _conv1(sd, f"{pre}conv1d_block.{i}.", x), bn_training))
        y = torch.relu(_bn(sd, pre + "recon_head.1.", _conv1(sd, pre + "recon_head.0.", x),
                           bn_training))
        res["recon"] = _conv1(sd, pre + "recon_head.3.", y)
        z = torch.cat([f_global[:, None, :].expand(-1, N, -1), x, xyz], dim=2)
        for i in (0, 3, 6):
            z = torch.relu(_bn(sd, f"{pre}face_head.{i + 1}.",
                               _conv1(sd, f"{pre}face_head.{i}.", z), bn_training))
        res["face"] = _conv1(sd, pre + "face_head.9.", z)
    return res


def pose_head(sd, pre, x_bnc, bn_training=False, dropout_p=0.0):
    """Rot_green / Rot_red / Pose_Ts forward (PoseR.py:26-39, PoseTs.py:31-45):
    two Conv-BN-ReLU, max over points, Conv-BN-ReLU, dropout, Conv."""
    x = torch.relu(_bn(sd, pre + "bn1.", _conv1(sd, pre + "conv1.", x_bnc), bn_training))
    x = torch.relu(_bn(sd, pre + "bn2.", _conv1(sd, pre + "conv2.", x), bn_training))
    x = x.amax(dim=1, keepdim=True)
    x = torch.relu(_bn(sd, pre + "bn3.", _conv1(sd, pre + "conv3.", x), bn_training))
    if dropout_p > 0:
        x = F.dropout(x, dropout_p, True)
    return _conv1(sd, pre + "conv4.", x).squeeze(1)


def posenet9d(sd, points, obj_id, k=20, S=7, train=True, bn_training=False, samples=None,
              rf_indices=None, dropout_p=0.0, pre="posenet."):
    """PoseNet9D.forward, PoseNet9D.py:23-52."""
    B, N, _ = points.shape
    mean = points.mean(dim=1, keepdim=True)
    centred = points - mean
    fr = face_recon(sd, centred, obj_id, k, S, train, bn_training, samples, rf_indices,
                    pre + "face_recon.")
    feat = fr["feat"]
    out = {"feat": feat, "rf_indices": fr["rf_indices"], "samples": fr["samples"]}
    if train:
        face = fr["face"]
        out["recon"] = fr["recon"] + mean
        fn = face[:, :, :18].view(B, N, 6, 3)
        out["face_normal"] = fn / torch.norm(fn, dim=-1, keepdim=True)
        out["face_dis"] = face[:, :, 18:24]
        out["face_f"] = torch.sigmoid(face[:, :, 24:])
    green = pose_head(sd, pre + "rot_green.", feat, bn_training, dropout_p)
    red = pose_head(sd, pre + "rot_red.", feat, bn_training, dropout_p)
    out["p_green_R"] = green[:, 1:] / (torch.norm(green[:, 1:], dim=1, keepdim=True) + 1e-6)
    out["p_red_R"] = red[:, 1:] / (torch.norm(red[:, 1:], dim=1, keepdim=True) + 1e-6)
    out["f_green_R"] = torch.sigmoid(green[:, 0])
    out["f_red_R"] = torch.sigmoid(red[:, 0])
    ts = pose_head(sd, pre + "ts.", torch.cat([feat, centred], dim=2), bn_training, dropout_p)
    out["Pred_T"] = ts[:, 0:3] + mean[:, 0, :]
    out["Pred_s"] = ts[:, 3:6]
    return out
