"""B200-native mirror of the reference's network/fs_net_repo/gcn3d.py.

Same public names, argument meaning, return shapes/dtypes and parameter
names/shapes/initialisers as the reference module (gcn3d.py:15-246), so
FaceRecon / PoseNet9D / HSPose and the published checkpoints work unchanged —
but nothing here materialises an (N,N) distance matrix or a (B,N,k,S*C)
neighbour tensor: every function is one or two launches of a hand-written
sm_100a kernel behind the C ABI (include/hspose_b200.h) via `ops`.

There is no CPU path: CPU tensors raise HSPoseLibraryError.
"""
import contextlib
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

# --------------------------------------------------------------------------
# Geometric-KNN memo.  One FaceRecon forward asks for the same xyz neighbour
# table up to 4 times per resolution (layer RF-P, ORL of two layers, pooling;
# reference call log: SURVEY.md Appendix B).  Inside `neighbor_cache()` the
# table is computed once per (tensor, k) and smaller k are served as prefixes
# of the ascending list (exact: our order is total on (distance, index)).
# --------------------------------------------------------------------------
_cache = None
_pool_rows_provider = None   # callable(vertice_num, pool_num, device) -> (pool_num,) int32 device rows
_forced_rf = None     # list of (B,N,k) index tensors consumed in order (teacher forcing, protocol T2)
_recorded_rf = None   # list receiving the RF-F tables a forward produced


@contextlib.contextmanager
def neighbor_cache():
    global _cache
    prev, _cache = _cache, {}
    try:
        yield
    finally:
        _cache = prev


def set_pool_rows_provider(fn):
    """Route Pool_layer's sample draw through `fn` (engine.TrainStep uses it to feed the
    permutation from a static device buffer so the whole step can live in a CUDA graph; the
    draw itself is still `torch.randperm(vertice_num)[:pool_num]` on the CPU generator)."""
    global _pool_rows_provider
    prev, _pool_rows_provider = _pool_rows_provider, fn
    return prev


@contextlib.contextmanager
def force_rf_indices(indices):
    """Teacher-force the feature-space neighbour tables (protocol T2)."""
    global _forced_rf
    prev, _forced_rf = _forced_rf, list(indices)
    try:
        yield
    finally:
        _forced_rf = prev


@contextlib.contextmanager
def record_rf_indices(out_list):
    global _recorded_rf
    prev, _recorded_rf = _recorded_rf, out_list
    try:
        yield out_list
    finally:
        _recorded_rf = prev


def _geo_index32(vertices, k):
    """(B,N,kk>=k) int32 ascending xyz neighbours (rank 0 dropped); returns (table, kk)."""
    if _cache is None:
        return ops.knn3(vertices, vertices, k)[1], k
    key = (vertices.data_ptr(), tuple(vertices.shape), vertices._version)
    hit = _cache.get(key)
    if hit is None or hit[1] < k:
        hit = (ops.knn3(vertices, vertices, k)[1], k, vertices)  # keep the tensor alive (data_ptr key)
        _cache[key] = hit
    return hit[0], hit[1]


def _geo_index32_exact(vertices, k):
    table, kk = _geo_index32(vertices, k)
    return table if kk == k else table[:, :, :k].contiguous()


def _checked_index(index, n_rows, what):
    """Caller-supplied neighbour / row indices are range-checked before a kernel dereferences them (PyTorch's
    own indexing raises for these; the kernels would read out of bounds).  One host sync — only on the API-parity
    functions and the teacher-forcing hook, never on the indices the library computes itself."""
    if index.numel() and (int(index.min()) < 0 or int(index.max()) >= n_rows):
        raise IndexError(f"{what}: index out of range [0, {n_rows})")
    return index


def _feature_index32(feature_map, k):
    if _forced_rf is not None:
        idx = _checked_index(_forced_rf.pop(0), feature_map.shape[1], "force_rf_indices")
        idx32 = idx.to(device=feature_map.device, dtype=torch.int32).contiguous()
    else:
        idx32 = ops.knn_feat(feature_map, k)[1]
    if _recorded_rf is not None:
        _recorded_rf.append(idx32)
    return idx32


# ------------------------------------------------------------------ functions
def get_neighbor_index(vertices: "(bs, vertice_num, D)", neighbor_num: int):
    """Reference gcn3d.py:15-24.  Return: (bs, vertice_num, neighbor_num) int64."""
    if vertices.shape[-1] == 3:
        return ops.knn3(vertices, vertices, neighbor_num, want64=True, want32=False)[0]
    return ops.knn_feat(vertices, neighbor_num, want64=True, want32=False)[0]


def get_nearest_index(target: "(bs, v1, 3)", source: "(bs, v2, 3)"):
    """Reference gcn3d.py:27-36.  Return: (bs, v1, 1) int64."""
    return ops.knn3(target, source, 1, drop_first=0, formula=ops.DIST_NEAREST, want64=True,
                    want32=False)[0]


def indexing_neighbor_new(tensor: "(bs, vertice_num, dim)", index: "(bs, v_out, neighbor_num)"):
    """Reference gcn3d.py:39-47 (materialising form, API parity only — the fused
    kernels never call it).  Return: (bs, v_out, neighbor_num, dim)."""
    bs, v_out, n = index.shape
    _checked_index(index, tensor.shape[1], "indexing_neighbor_new")
    rows = ops.gather_rows(tensor, index.reshape(bs, v_out * n).to(torch.int32))
    return rows.view(bs, v_out, n, tensor.shape[2])


def get_neighbor_direction_norm(vertices, neighbor_index, return_unnormed=False):
    """Reference gcn3d.py:49-59.  Return: (bs, vertice_num, neighbor_num, 3) fp32."""
    _checked_index(neighbor_index, vertices.shape[1], "get_neighbor_direction_norm")
    return ops.direction_norm(vertices, neighbor_index.to(torch.int32), return_unnormed)


def get_receptive_fields(neighbor_num, vertices, feature_map=None, mode='RF-F'):
    """Reference gcn3d.py:189-209."""
    assert mode in ['RF-F', 'RF-P']
    if mode == 'RF-F':
        assert feature_map is not None, "The feature_map should be provided if 'RF-F' is used"
        feat = feature_map
    else:
        feat = vertices
    neighbor_index = get_neighbor_index(feat, neighbor_num)
    return get_neighbor_direction_norm(vertices, neighbor_index), neighbor_index


_ORL_ONE_NODE = os.environ.get("HSP_ORL_ONE_NODE", "1") != "0"    # A/B switch (tools/, tests)


def get_ORL_global(feature, vertices, neighbor_num):
    """Reference gcn3d.py:211-218.  Return: (bs, vertice_num, C) (per-object constant, repeated)."""
    G = ops.orl_global(feature, _geo_index32_exact(vertices, neighbor_num))
    return G.unsqueeze(1).repeat(1, feature.size(1), 1)


def _orl_fuse(feature, vertices, neighbor_num, conv2_weight, f_STE=None, ste_xyz_weight=None):
    """ORL_forward (gcn3d.py:109-113 / :183-187) without the cat/repeat:
    conv2(cat[f, G]) = f @ W2[:, :C]^T + (G @ W2[:, C:]^T) broadcast over points, and the
    layer's `+ f_STE` (gcn3d.py:90 / :156) folded into the same pass (K5d) when given."""
    C = feature.shape[2]
    W2 = conv2_weight.squeeze(-1)
    if (torch.is_autocast_enabled("cuda") and feature.is_cuda and C % 8 == 0 and feature.dtype == torch.float32
            and _ORL_ONE_NODE):
        # one autograd node: ORL + both 1x1 GEMMs + the residual pass; backward without fill / add passes
        if ste_xyz_weight is not None:
            return ops.orl_fuse(feature, _geo_index32_exact(vertices, neighbor_num), W2, None, vertices.float(),
                                ste_xyz_weight.float())
        return ops.orl_fuse(feature, _geo_index32_exact(vertices, neighbor_num), W2, f_STE)
    G = ops.orl_global(feature, _geo_index32_exact(vertices, neighbor_num))  # (B,C)
    if torch.is_autocast_enabled("cuda") and feature.is_cuda and C % 8 == 0:
        W2f, W2g = ops.split_halves(W2, C)                                   # one cat in the backward
        lin = ops.linear_tc(feature, W2f)                                    # K6 (bf16 operands, fp32 accumulate)
        gproj = ops.linear_tc(G, W2g).float()                                # (B,C)
    else:
        lin = F.linear(feature, W2[:, :C])
        with torch.autocast("cuda", enabled=False):
            gproj = F.linear(G.float(), W2[:, C:].float())                   # (B,C)
    if C % 4 == 0 and feature.dtype == torch.float32:
        if ste_xyz_weight is not None:   # surface layer: STE = 3 -> C linear on xyz, evaluated in the same pass
            return ops.residual_sum(feature, lin, gproj, None, vertices.float(), ste_xyz_weight.float())
        return ops.residual_sum(feature, lin, gproj, f_STE)
    if ste_xyz_weight is not None:
        with torch.autocast("cuda", enabled=False):
            f_STE = F.linear(vertices.float(), ste_xyz_weight.float())
    out = feature + lin + gproj.unsqueeze(1)
    return out if f_STE is None else out + f_STE


def _unit_dirs(directions):
    """F.normalize(directions, dim=0) (reference gcn3d.py:95, :162); one launch each way on the mixed-precision
    train path, the PyTorch composite (bit-for-bit the reference's) on the fp32 parity path."""
    if directions.is_cuda and torch.is_autocast_enabled("cuda"):
        return ops.normalize_dirs(directions)
    return F.normalize(directions, dim=0)


# -------------------------------------------------------------------- layers
class HSlayer_surface(nn.Module):
    """Reference gcn3d.py:61-113 (same parameters: directions (3,S*C),
    STE_layer.weight (C,3,1), conv2.weight (C,2C,1))."""

    def __init__(self, kernel_num, support_num):
        super().__init__()
        self.feat_k = 8
        self.kernel_num = kernel_num
        self.support_num = support_num
        self.relu = nn.ReLU(inplace=True)
        self.directions = nn.Parameter(torch.FloatTensor(3, support_num * kernel_num))
        self.STE_layer = nn.Conv1d(3, kernel_num, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(2 * kernel_num, kernel_num, kernel_size=1, bias=False)
        self.initialize()

    def initialize(self):
        stdv = 1. / math.sqrt(self.support_num * self.kernel_num)
        self.directions.data.uniform_(-stdv, stdv)

    def forward(self, vertices: "(bs, vertice_num, 3)", neighbor_num: 'int'):
        # STE (Conv1d 3 -> C on the coordinates, gcn3d.py:86) is folded into the fused residual pass
        feature = self.graph_conv(None, vertices, neighbor_num)
        return _orl_fuse(feature, vertices, neighbor_num, self.conv2.weight,
                         ste_xyz_weight=self.STE_layer.weight.squeeze(-1))

    def graph_conv(self, receptive_fields_norm, vertices, neighbor_num):
        """K3.  `receptive_fields_norm` is accepted for signature parity and ignored:
        the unit directions are recomputed in-kernel from the RF-P neighbour table."""
        idx32 = _geo_index32_exact(vertices, neighbor_num)
        dirn = _unit_dirs(self.directions)
        return ops.surface_conv(vertices, idx32, dirn, self.support_num, self.kernel_num)

    def ORL_forward(self, feature, vertices, neighbor_num):
        return _orl_fuse(feature, vertices, neighbor_num, self.conv2.weight)


class HS_layer(nn.Module):
    """Reference gcn3d.py:116-187 (parameters: weights (Cin,(S+1)Cout), bias,
    directions (3,S*Cout), STE_layer.weight (Cout,Cin,1), conv2.weight (Cout,2Cout,1))."""

    def __init__(self, in_channel, out_channel, support_num):
        super().__init__()
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.support_num = support_num
        self.relu = nn.ReLU(inplace=True)
        self.weights = nn.Parameter(torch.FloatTensor(in_channel, (support_num + 1) * out_channel))
        self.bias = nn.Parameter(torch.FloatTensor((support_num + 1) * out_channel))
        self.directions = nn.Parameter(torch.FloatTensor(3, support_num * out_channel))
        self.feat_k = 8
        self.STE_layer = nn.Conv1d(self.in_channel, self.out_channel, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(2 * out_channel, out_channel, kernel_size=1, bias=False)
        self.initialize()

    def initialize(self):
        stdv = 1. / math.sqrt(self.out_channel * (self.support_num + 1))
        self.weights.data.uniform_(-stdv, stdv)
        self.bias.data.uniform_(-stdv, stdv)
        self.directions.data.uniform_(-stdv, stdv)

    def forward(self, vertices, feature_map, neighbor_num):
        if torch.is_autocast_enabled("cuda") and feature_map.is_cuda and self.in_channel % 8 == 0:
            f_STE = ops.linear_tc(feature_map, self.STE_layer.weight.squeeze(-1))    # K6
        else:
            f_STE = F.linear(feature_map, self.STE_layer.weight.squeeze(-1))
        neighbor_index = _feature_index32(feature_map, neighbor_num)          # RF-F (K2)
        feature = self.graph_conv(None, neighbor_index, feature_map, vertices, neighbor_num)
        return _orl_fuse(feature, vertices, neighbor_num, self.conv2.weight, f_STE)

    def graph_conv(self, receptive_fields_norm, neighbor_index, feature_map, vertices,
                   neighbor_num):
        """K4.  fm @ W + b is the dense contraction (library GEMM); the gather of the
        support rows, theta, the product, max over neighbours, mean over supports and the
        centre term are one kernel."""
        idx32 = neighbor_index if neighbor_index.dtype == torch.int32 else neighbor_index.to(torch.int32)
        if torch.is_autocast_enabled("cuda"):
            # mixed precision: bf16 P straight from the tensor-core GEMM into the gather kernel
            return ops.hs_conv_mixed(vertices, idx32, _unit_dirs(self.directions), feature_map,
                                     self.weights, self.bias, self.support_num, self.out_channel)
        P = torch.addmm(self.bias, feature_map.reshape(-1, self.in_channel), self.weights)
        P = P.view(feature_map.shape[0], feature_map.shape[1], -1)
        dirn = F.normalize(self.directions, dim=0)
        return ops.graph_conv(vertices, idx32, dirn, P, self.support_num, self.out_channel)

    def ORL_forward(self, feature_fuse, vertices, neighbor_num):
        return _orl_fuse(feature_fuse, vertices, neighbor_num, self.conv2.weight)


class Pool_layer(nn.Module):
    """Reference gcn3d.py:220-246.  The sample is drawn from the CPU generator with the
    same call (`torch.randperm(vertice_num)[:pool_num]`, :243) so RNG streams stay aligned."""

    def __init__(self, pooling_rate: int = 4, neighbor_num: int = 4):
        super().__init__()
        self.pooling_rate = pooling_rate
        self.neighbor_num = neighbor_num

    def forward(self, vertices, feature_map):
        bs, vertice_num, _ = vertices.size()
        table, kk = _geo_index32(vertices, self.neighbor_num)
        pool_num = int(vertice_num / self.pooling_rate)
        if _pool_rows_provider is not None:
            rows = _pool_rows_provider(vertice_num, pool_num, vertices.device)
        else:
            sample_idx = torch.randperm(vertice_num)[:pool_num]
            rows = sample_idx.to(device=vertices.device, dtype=torch.int32, non_blocking=True)
        vertices_pool = ops.gather_rows(vertices, rows.unsqueeze(0).expand(bs, -1).contiguous())
        # max over the 4 nearest evaluated ONLY at the sampled rows (K5b)
        feature_map_pool = ops.gather_max(feature_map, table, rows, kuse=self.neighbor_num)
        return vertices_pool, feature_map_pool
