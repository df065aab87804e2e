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


def _rf_recorder(gcn3d):
    """Wrap the reference's get_receptive_fields to log the RF-F index tensors."""
    log = []
    orig = gcn3d.get_receptive_fields

    def wrapped(neighbor_num, vertices, feature_map=None, mode='RF-F'):
        d, i = orig(neighbor_num, vertices, feature_map=feature_map, mode=mode)
        if mode == 'RF-F':
            log.append(i.clone())
        return d, i
    gcn3d.get_receptive_fields = wrapped
    return log, lambda: setattr(gcn3d, "get_receptive_fields", orig)


def golden_e2e_eval(FLAGS, gcn3d):
    """PoseNet9D built and run as evaluation/evaluate.py does (FLAGS.train=0 before
    construction, .eval()), deterministic weights (oracle.synth.fill_params)."""
    sys.path.insert(0, ROOT)
    from hspose_b200.synth import fill_params, synth_batch
    from network.fs_net_repo.PoseNet9D import PoseNet9D
    FLAGS.train = 0
    arrays = {}
    for k in (20, 16):
        FLAGS.gcn_n_num = k
        net = fill_params(PoseNet9D()).eval()
        batch = synth_batch(2, 1028, seed=1, train=False)
        log, undo = _rf_recorder(gcn3d)
        torch.manual_seed(1234)
        with torch.no_grad():
            out = net(batch["PC"], batch["obj_id"])
        undo()
        torch.manual_seed(1234)
        s1 = torch.randperm(1028)[:257]
        s2 = torch.randperm(257)[:64]
        names = ["p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s"]
        for n, t in zip(names, out[4:]):
            arrays[f"k{k}_{n}"] = t
        for i, r in enumerate(log):
            arrays[f"k{k}_rf{i}"] = small_idx(r)
        arrays[f"k{k}_sample1"] = small_idx(s1)
        arrays[f"k{k}_sample2"] = small_idx(s2)
        # backbone feature (concat) on a strided subset of points, via a forward hook
        feats = []
        h = net.face_recon.register_forward_hook(lambda m, i, o: feats.append(o[2]))
        torch.manual_seed(1234)
        with torch.no_grad():
            net(batch["PC"], batch["obj_id"])
        h.remove()
        arrays[f"k{k}_feat_s16"] = feats[0][:, ::16, :].contiguous()
    arrays["n_state_keys"] = np.int64(len(net.state_dict()))
    FLAGS.gcn_n_num = 20
    FLAGS.train = 1
    save("e2e_eval", **arrays)


def golden_e2e_eval_b16(FLAGS, gcn3d):
    """BASELINE.json configs[1] as written: batch = 16, N = 1028, fp32 forward only, KNN k = 16 (and the
    shipped k = 20): the six pose / size outputs and the RF-F neighbour tables of the reference run."""
    sys.path.insert(0, ROOT)
    from hspose_b200.synth import fill_params, synth_batch
    from network.fs_net_repo.PoseNet9D import PoseNet9D
    FLAGS.train = 0
    arrays = {}
    for k in (16, 20):
        FLAGS.gcn_n_num = k
        net = fill_params(PoseNet9D()).eval()
        batch = synth_batch(16, 1028, seed=1, train=False)
        log, undo = _rf_recorder(gcn3d)
        torch.manual_seed(1234)
        with torch.no_grad():
            out = net(batch["PC"], batch["obj_id"])
        undo()
        for n, t in zip(["p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s"], out[4:]):
            arrays[f"k{k}_{n}"] = t
        for i, r in enumerate(log):
            arrays[f"k{k}_rf{i}"] = small_idx(r)
    FLAGS.gcn_n_num = 20
    FLAGS.train = 1
    save("e2e_eval_b16", **arrays)


def golden_e2e_train(FLAGS, gcn3d):
    """HSPose('PoseNet_only') train-mode step as engine/train.py:76-107 runs it: forward with
    do_loss=True, sum of all loss terms, backward.  Dropout p=0 and augmentation
    probabilities 0 (device RNG streams differ between CPU and GPU)."""
    sys.path.insert(0, ROOT)
    from hspose_b200.synth import fill_params, synth_batch
    from network.HSPose import HSPose
    FLAGS.train = 1
    for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro"):
        setattr(FLAGS, n, 0.0)
    net = fill_params(HSPose("PoseNet_only")).train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    batch = synth_batch(4, 1028, seed=2, train=True)
    log, undo = _rf_recorder(gcn3d)
    torch.manual_seed(4321)
    out, losses = net(**batch, do_loss=True)
    undo()
    arrays = {}
    fs = {k: v.reshape(()) for k, v in losses["fsnet_loss"].items()}
    total_fs = sum(fs.values())
    for k, v in fs.items():
        arrays["loss_fs_" + k] = v
    for grp in ("recon_loss", "geo_loss", "prop_loss"):
        for k, v in losses[grp].items():
            arrays[f"loss_{grp}_{k}"] = v.reshape(())
    total_fs.backward()
    for i, r in enumerate(log):
        arrays[f"rf{i}"] = small_idx(r)
    for n in ("recon", "face_normal", "face_dis", "face_f"):
        arrays["out_" + n] = out[n][:, ::64].contiguous()
    for n in ("p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s"):
        arrays["out_" + n] = out[n]
    gn = {}
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        gn[name] = float(p.grad.norm())
        if p.numel() <= 4096 or name.endswith("directions"):
            arrays["grad::" + name] = p.grad
    arrays["grad_norm_names"] = np.array(sorted(gn.keys()))
    arrays["grad_norm_values"] = np.array([gn[k] for k in sorted(gn.keys())], dtype=np.float64)
    # BN running stats after the step (momentum update) for two layers
    sd = net.state_dict()
    for n in ("posenet.face_recon.bn1.running_mean", "posenet.face_recon.bn1.running_var",
              "posenet.rot_green.bn2.running_mean"):
        arrays["post::" + n] = sd[n]
    for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro"):
        setattr(FLAGS, n, {"aug_pc_pro": 0.2}.get(n, 0.3))
    save("e2e_train", **arrays)


def golden_augment(FLAGS):
    """HSPose.data_augment of the reference (network/HSPose.py:185-256 over
    datasets/data_augmentation.py:70-190) on CPU: once with every branch forced on (all
    probabilities 1) and once with the shipped probabilities; RNG seeded before each call."""
    sys.path.insert(0, ROOT)
    from hspose_b200.synth import synth_batch
    from network.HSPose import HSPose
    FLAGS.train = 1
    net = HSPose("PoseNet_only")
    batch = synth_batch(8, 1028, seed=5, train=True)
    arrays = {}
    keep = {n: getattr(FLAGS, n) for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro")}
    for tag, probs in (("all", dict(aug_pc_pro=1.0, aug_rt_pro=1.0, aug_bb_pro=1.0, aug_bc_pro=1.0)),
                       ("default", dict(aug_pc_pro=0.2, aug_rt_pro=0.3, aug_bb_pro=0.3, aug_bc_pro=0.3))):
        for n, v in probs.items():
            setattr(FLAGS, n, v)
        torch.manual_seed(777)
        with torch.no_grad():
            PC, R, t, s = net.data_augment(batch["PC"].clone(), batch["gt_R"].clone(), batch["gt_t"].clone(),
                                           batch["gt_s"].clone(), batch["mean_shape"], batch["sym"],
                                           batch["aug_bb"], batch["aug_rt_t"], batch["aug_rt_r"],
                                           batch["model_point"].clone(), batch["nocs_scale"], batch["obj_id"])
        arrays.update({f"{tag}_PC": PC, f"{tag}_R": R, f"{tag}_t": t, f"{tag}_s": s})
    for n, v in keep.items():
        setattr(FLAGS, n, v)
    save("aug", **arrays)


def golden_losses(FLAGS):
    """The reference's recon_6face / geo / prop loss modules (losses/*.py) on synthetic
    predictions: every term and the gradient of each group's sum w.r.t. every prediction."""
    from losses.geometry_loss import geo_transform_loss
    from losses.prop_loss import prop_rot_loss
    from losses.recon_loss import recon_6face_loss
    from tools.geom_utils import generate_RT
    sys.path.insert(0, ROOT)
    from hspose_b200.synth import synth_predictions
    pred, gt = synth_predictions(12, 257, seed=11)
    leaves = {k: v.clone().requires_grad_() for k, v in pred.items()}
    arrays = {}
    groups = {
        "recon": lambda p: recon_6face_loss()(
            ['Per_point', 'Point_voting'],
            {'F_n': p["face_normal"], 'F_d': p["face_dis"], 'F_c': p["face_f"], 'Rot1': p["p_green_R"],
             'Rot1_f': p["f_green_R"].detach(), 'Rot2': p["p_red_R"], 'Rot2_f': p["f_red_R"].detach(),
             'Tran': p["Pred_T"], 'Size': p["Pred_s"]},
            {'R': gt["gt_R"], 'T': gt["gt_t"], 'Size': gt["gt_s"], 'Mean_shape': gt["mean_shape"],
             'Points': gt["PC"]}, gt["sym"], gt["obj_id"]),
        "geo": lambda p: geo_transform_loss()(
            ['Geo_point'],
            {'Rot1': p["p_green_R"], 'Rot2': p["p_red_R"], 'Tran': p["Pred_T"], 'Size': p["Pred_s"],
             'Rot1_f': p["f_green_R"].detach(), 'Rot2_f': p["f_red_R"].detach()},
            {'Points': gt["PC"], 'R': gt["gt_R"], 'T': gt["gt_t"], 'Mean_shape': gt["mean_shape"]}, gt["sym"]),
        "prop": lambda p: prop_rot_loss()(
            ['Prop_pm', 'Prop_sym'],
            {'Recon': p["recon"], 'Rot1': p["p_green_R"], 'Rot2': p["p_red_R"], 'Tran': p["Pred_T"],
             'Scale': p["Pred_s"], 'Rot1_f': p["f_green_R"].detach(), 'Rot2_f': p["f_red_R"].detach()},
            {'Points': gt["PC"], 'R': gt["gt_R"], 'T': gt["gt_t"], 'Mean_shape': gt["mean_shape"]}, gt["sym"]),
    }
    for gname, fn in groups.items():
        for v in leaves.values():
            v.grad = None
        terms = fn(leaves)
        total = 0.0
        for k, v in terms.items():
            if torch.is_tensor(v):
                arrays[f"{gname}::{k}"] = v.detach().reshape(())
                total = total + v
        total.backward()
        for k, v in leaves.items():
            if v.grad is not None and float(v.grad.abs().max()) > 0:
                arrays[f"{gname}::grad::{k}"] = v.grad.clone()
    # evaluation post-processing (tools/geom_utils.py:232-244)
    with torch.no_grad():
        arrays["generate_RT"] = generate_RT([pred["p_green_R"], pred["p_red_R"]],
                                            [pred["f_green_R"], pred["f_red_R"]], pred["Pred_T"], "vec", gt["sym"])
    save("losses", **arrays)


def golden_optim():
    """The reference's Ranger (tools/torch_utils/solver/ranger2020.py) preceded by
    clip_grad_norm_(.., 5) as engine/train.py:105-110 runs it, 8 steps (lookahead fires at 6) on a
    small parameter set with seeded gradients; parameters saved after steps 1, 5, 6 and 8."""
    from tools.torch_utils.solver.ranger2020 import Ranger
    shapes = [(8, 16), (16,), (4, 8, 1), (3, 24), (5,), (32, 40)]
    g = torch.Generator().manual_seed(21)
    params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.3) for s in shapes]
    arrays = {f"p0_{i}": p.detach().clone() for i, p in enumerate(params)}
    opt = Ranger(params, lr=1e-2)
    for step in range(1, 9):
        scale = 30.0 if step in (2, 7) else 0.5          # steps 2 and 7 exceed the clip norm of 5
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=g) * scale
            arrays[f"g{step}_{i}"] = p.grad.clone()
        arrays[f"norm{step}"] = torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()
        if step in (1, 5, 6, 8):
            for i, p in enumerate(params):
                arrays[f"p{step}_{i}"] = p.detach().clone()
    save("optim", **arrays)


def golden_prestage(FLAGS):
    """The reference's input pre-stage on synthetic depth ROIs: `PoseDataset._depth_to_pcl` + `/1000` +
    `_sample_points` (datasets/load_data.py:277,307-333) and `PC_sample` (network/point_sample/pc_sample.py).
    Shim: numpy >= 1.24 removed `np.float` (load_data.py:325 uses it); it is aliased to `float` here."""
    np.float = float
    from datasets.load_data import PoseDataset
    from network.point_sample.pc_sample import PC_sample
    rng = np.random.RandomState(5)
    B, H, W = 4, 48, 64
    n = int(FLAGS.random_points)
    # depth in millimetres (uint16-valued), holes (0), an object mask of different sizes per object:
    # object 0: > n valid pixels, 1: < n (tile / replace=True), 2: exactly structured, 3: few pixels
    depth = (rng.randint(400, 1500, size=(B, H, W))).astype(np.float32)
    depth[rng.rand(B, H, W) < 0.15] = 0.0
    mask = np.zeros((B, H, W), dtype=np.float32)
    mask[0, 4:44, 6:60] = 1.0
    mask[1, 10:30, 12:40] = 1.0
    mask[2, :, :] = (rng.rand(H, W) < 0.6)
    mask[3, 20:23, 30:37] = 1.0
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    xymap = np.stack([xs * 2.5 + 100.25, ys * 2.5 + 60.5], 0).astype(np.float32)
    xymap = np.repeat(xymap[None], B, 0) + rng.rand(B, 1, 1, 1).astype(np.float32) * 40
    K = np.array([[577.5, 0, 319.5], [0, 577.5, 239.5], [0, 0, 1.0]])
    camK = np.repeat(K[None], B, 0) + rng.rand(B, 3, 3) * np.array([[3.0, 0, 3.0], [0, 3.0, 3.0], [0, 0, 0]])
    out = {"depth": depth, "mask": mask, "xymap": xymap, "camK64": camK, "n_pts": np.int64(n)}
    # (a) the loader's functions (float64 arithmetic)
    np.random.seed(123)
    for b in range(B):
        pcl = PoseDataset._depth_to_pcl(None, depth[b], camK[b], xymap[b], mask[b]) / 1000.0
        out[f"pcl_{b}"] = pcl
        total = pcl.shape[0]
        if total > n:      # reproduce the draw _sample_points makes, and record it
            st = np.random.get_state()
            ids = np.random.permutation(total)[:n]
            np.random.set_state(st)
            out[f"ids_{b}"] = ids.astype(np.int32)
        out[f"sampled_{b}"] = PoseDataset._sample_points(None, pcl, n)
        assert out[f"sampled_{b}"].dtype == np.float32
    # (b) PC_sample (torch float32), objects 0..2 (object 3 has > 1 pixel too, keep all four)
    FLAGS.sample_method = "basic"
    np.random.seed(321)
    st = np.random.get_state()
    PC = PC_sample(torch.from_numpy(mask)[:, None], torch.from_numpy(depth)[:, None],
                   torch.from_numpy(camK.astype(np.float32)), torch.from_numpy(xymap))
    out["PC_sample"] = PC
    np.random.set_state(st)
    chooses = []
    for b in range(B):
        l_all = int(((mask[b] * (depth[b] > 0)) > 0).sum())
        chooses.append(np.random.choice(l_all, n, replace=l_all < n))
    out["PC_choose"] = np.stack(chooses).astype(np.int32)
    save("prestage", **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    _args = sys.argv[1:]
    FLAGS, gcn3d = import_reference()
    which = _args or ["knn", "ops", "e2e_eval", "e2e_train", "aug", "losses", "optim", "e2e_eval_b16", "prestage"]
    if "knn" in which:
        golden_knn(gcn3d)
    if "ops" in which:
        golden_ops(gcn3d)
    if "e2e_eval" in which:
        golden_e2e_eval(FLAGS, gcn3d)
    if "e2e_eval_b16" in which:
        golden_e2e_eval_b16(FLAGS, gcn3d)
    if "e2e_train" in which:
        golden_e2e_train(FLAGS, gcn3d)
    if "aug" in which:
        golden_augment(FLAGS)
    if "losses" in which:
        golden_losses(FLAGS)
    if "optim" in which:
        golden_optim()
    if "prestage" in which:
        golden_prestage(FLAGS)
