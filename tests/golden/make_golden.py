#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (/root/reference).

Run in the build container only (the reference tree does not travel to the GPU
box):   python tests/golden/make_golden.py
The committed .npz files pin both oracle/ (CPU restatement) and the CUDA path.
Nothing here is imported by the product.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HSPOSE_REFERENCE", "/root/reference")


def import_reference(train=1):
    """Import the reference package tree with its unused heavy deps stubbed
    (matplotlib, mmcv, detectron2, termcolor — SURVEY.md Appendix D)."""
    sys.path.insert(0, REF)
    sys.argv = ["x"]

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("matplotlib")
    stub("matplotlib.pyplot", axis=None)
    stub("mmcv", Config=dict)
    stub("detectron2")
    stub("detectron2.config", CfgNode=dict)
    stub("detectron2.solver", WarmupCosineLR=None, WarmupMultiStepLR=None)
    stub("termcolor", colored=lambda s, *a, **k: s)
    import absl.flags as flags
    import config.config  # noqa: F401  (defines the flags)
    if not flags.FLAGS.is_parsed():
        flags.FLAGS(sys.argv)
    flags.FLAGS.train = train
    import network.fs_net_repo.gcn3d as gcn3d
    return flags.FLAGS, gcn3d


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def small_idx(t):
    return t.to(torch.int16)


def golden_knn(gcn3d):
    sys.path.insert(0, ROOT)
    from oracle.synth import tiled_cloud
    g = torch.Generator().manual_seed(11)
    arrays = {}
    # centred clouds — what FaceRecon feeds the KNN (PoseNet9D.py:25)
    v = torch.randn(2, 300, 3, generator=g) * 0.05
    arrays["c_xyz"] = v
    for k in (4, 8, 20, 32):
        arrays[f"c_idx_k{k}"] = small_idx(gcn3d.get_neighbor_index(v, k))
    # un-centred (+0.8 m): distances are heavily quantised -> exact ties
    u = v + torch.tensor([0.0, 0.0, 0.8])
    arrays["u_xyz"] = u
    arrays["u_idx_k20"] = small_idx(gcn3d.get_neighbor_index(u, 20))
    # pooled sizes used by the backbone (257 -> k=20, 64 -> k=8)
    w = torch.randn(2, 64, 3, generator=g) * 0.05
    arrays["p_xyz"] = w
    arrays["p_idx_k8"] = small_idx(gcn3d.get_neighbor_index(w, 8))
    # nearest (up-sampling): target 300 pts, source = subset of 75 / 18
    for m in (75, 18):
        sel = torch.randperm(300, generator=g)[:m]
        src = v[:, sel].contiguous()
        arrays[f"n_src{m}"] = src
        arrays[f"n_idx{m}"] = small_idx(gcn3d.get_nearest_index(v, src))
    # duplicates: 120 unique points tiled to 300
    t = tiled_cloud(2, 300, 120, seed=5)
    arrays["t_xyz"] = t
    arrays["t_idx_k20"] = small_idx(gcn3d.get_neighbor_index(t, 20))
    # feature space, D=128 and 256, post-ReLU statistics with a large common mean
    f = torch.relu(torch.randn(2, 257, 128, generator=g) + 1.0)
    arrays["f128"] = f
    arrays["f128_idx_k20"] = small_idx(gcn3d.get_neighbor_index(f, 20))
    f2 = torch.relu(torch.randn(1, 64, 256, generator=g) + 1.0)
    arrays["f256"] = f2
    arrays["f256_idx_k8"] = small_idx(gcn3d.get_neighbor_index(f2, 8))
    save("knn", **arrays)


def golden_ops(gcn3d):
    """Function-level vectors for the fused ops (small channel counts)."""
    g = torch.Generator().manual_seed(23)
    B, N, k, S, C = 2, 96, 8, 7, 16
    xyz = torch.randn(B, N, 3, generator=g) * 0.05
    idx = gcn3d.get_neighbor_index(xyz, k)
    dirs = gcn3d.get_neighbor_direction_norm(xyz, idx)
    arrays = {"xyz": xyz, "idx": small_idx(idx), "dir_norm": dirs}

    # surface conv (HSlayer_surface.graph_conv, gcn3d.py:92-107) with free-standing params
    layer = gcn3d.HSlayer_surface(kernel_num=C, support_num=S)
    layer.directions.data = (torch.rand(3, S * C, generator=g) * 2 - 1) * 0.3
    arrays["surf_directions"] = layer.directions.data.clone()
    arrays["surf_out"] = layer.graph_conv(dirs, xyz, k)

    # HS conv (HS_layer.graph_conv, gcn3d.py:158-181), feature-space neighbours
    Cin = 24
    hs = gcn3d.HS_layer(Cin, C, support_num=S)
    fm = torch.relu(torch.randn(B, N, Cin, generator=g))
    hs.weights.data = (torch.rand(Cin, (S + 1) * C, generator=g) * 2 - 1) * 0.2
    hs.bias.data = (torch.rand((S + 1) * C, generator=g) * 2 - 1) * 0.1
    hs.directions.data = (torch.rand(3, S * C, generator=g) * 2 - 1) * 0.3
    rf_dirs, rf_idx = gcn3d.get_receptive_fields(k, xyz, feature_map=fm, mode="RF-F")
    arrays.update(hs_fm=fm, hs_weights=hs.weights.data.clone(), hs_bias=hs.bias.data.clone(),
                  hs_directions=hs.directions.data.clone(), hs_rf_idx=small_idx(rf_idx),
                  hs_out=hs.graph_conv(rf_dirs, rf_idx, fm, xyz, k))

    # ORL global (gcn3d.py:211-218) and pooling (gcn3d.py:220-246)
    feat = torch.randn(B, N, C, generator=g)
    arrays["orl_feat"] = feat
    arrays["orl_global"] = gcn3d.get_ORL_global(feat, xyz, k)[:, 0, :]
    torch.manual_seed(99)
    pool = gcn3d.Pool_layer(pooling_rate=4, neighbor_num=4)
    vp, fp = pool(xyz, feat)
    torch.manual_seed(99)
    arrays["pool_sample"] = small_idx(torch.randperm(N)[: N // 4])
    arrays["pool_xyz"] = vp
    arrays["pool_feat"] = fp
    # nearest up-sampling (FaceRecon.py:100-104)
    nn_idx = gcn3d.get_nearest_index(xyz, vp)
    arrays["up_idx"] = small_idx(nn_idx)
    arrays["up_out"] = gcn3d.indexing_neighbor_new(fp, nn_idx).squeeze(2)
    save("ops", **arrays)


if __name__ == "__main__":
    torch.set_num_threads(8)
    FLAGS, gcn3d = import_reference()
    which = sys.argv[1:] or ["knn", "ops"]
    if "knn" in which:
        golden_knn(gcn3d)
    if "ops" in which:
        golden_ops(gcn3d)
