"""GPU parity (T2 / T3, model level): the drop-in modules (gcn3d / FaceRecon /
PoseNet9D / HSPose of hs-pose_b200) against end-to-end golden vectors produced by
the REAL reference on identical inputs and weights (tests/golden/e2e_*.npz).

T2 = feature-space neighbour tables teacher-forced from the reference run:
     pose/size within 1e-5 (north_star tolerance, fp32).
T3 = free-running: reported against the reference's own noise floor
     (1e-3 on rotations when only the GEMM blocking changes).
"""
import numpy as np
import pytest
import torch

from oracle.synth import fill_params, synth_batch

pytestmark = pytest.mark.gpu
NAMES = ["p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s"]


@pytest.fixture(autouse=True)
def _fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _flags():
    import hspose_b200.flags as hf
    return hf.get_flags()


def _posenet(cuda, train, k=20):
    from hspose_b200.PoseNet9D import PoseNet9D
    F = _flags()
    F.train, F.gcn_n_num = train, k
    m = fill_params(PoseNet9D()).to(cuda)
    return m


def _run_eval(cuda, golden, k, forced, bs=2):
    from hspose_b200 import gcn3d
    g = golden("e2e_eval" if bs == 2 else f"e2e_eval_b{bs}")
    F = _flags()
    net = _posenet(cuda, 0, k).eval()
    try:
        batch = synth_batch(bs, 1028, seed=1, train=False)
        rec = []
        torch.manual_seed(1234)
        with torch.no_grad(), gcn3d.record_rf_indices(rec):
            if forced:
                rf = [torch.from_numpy(g[f"k{k}_rf{i}"].astype(np.int64)) for i in range(4)]
                with gcn3d.force_rf_indices(rf):
                    out = net(batch["PC"].to(cuda), batch["obj_id"].to(cuda))
            else:
                out = net(batch["PC"].to(cuda), batch["obj_id"].to(cuda))
    finally:
        F.train, F.gcn_n_num = 1, 20
    return g, dict(zip(NAMES, [t.cpu().numpy() for t in out[4:]])), rec


@pytest.mark.parametrize("bs", [2, 16])      # 16 = BASELINE.json configs[1] as written (batch 16, k = 16)
@pytest.mark.parametrize("k", [20, 16])
def test_posenet_eval_teacher_forced_1e5(cuda, golden, k, bs):
    g, out, _ = _run_eval(cuda, golden, k, forced=True, bs=bs)
    for n in NAMES:
        np.testing.assert_allclose(out[n], g[f"k{k}_{n}"], atol=1e-5, err_msg=n)


def test_backbone_feat_teacher_forced(cuda, golden):
    from hspose_b200 import gcn3d
    g = golden("e2e_eval")
    F = _flags()
    net = _posenet(cuda, 0, 20).eval()
    F.train = 0
    try:
        batch = synth_batch(2, 1028, seed=1, train=False)
        pc = batch["PC"].to(cuda)
        rf = [torch.from_numpy(g[f"k20_rf{i}"].astype(np.int64)) for i in range(4)]
        torch.manual_seed(1234)
        with torch.no_grad(), gcn3d.force_rf_indices(rf):
            _, _, feat = net.face_recon(pc - pc.mean(dim=1, keepdim=True), batch["obj_id"].to(cuda))
    finally:
        F.train = 1
    np.testing.assert_allclose(feat[:, ::16].cpu().numpy(), g["k20_feat_s16"], atol=2e-5)


def test_posenet_eval_free_running_report(cuda, golden, capsys):
    g, out, rec = _run_eval(cuda, golden, 20, forced=False)
    flips, rows = 0, 0
    for i in range(4):
        mine = np.sort(rec[i].cpu().numpy(), -1)
        ref = np.sort(g[f"k20_rf{i}"].astype(np.int32), -1)
        flips += int((mine != ref).any(-1).sum())
        rows += mine.shape[0] * mine.shape[1]
    diffs = {n: float(np.abs(out[n] - g[f"k20_{n}"]).max()) for n in NAMES}
    with capsys.disabled():
        print(f"\n[T3 free-running] RF-F neighbour-set flips {flips}/{rows}; max-abs diffs {diffs}")
    # Flips cascade through the 4 RF-F layers (the reference's own self-noise floor is ~2e-3 on the rotations when
    # only its GEMM blocking changes).  Measured on the B200 (deterministic): 190 / 3212 rows = 5.9 %,
    # p_green_R 2.6e-4, p_red_R 4.2e-4, f_* 1e-6 .. 5e-6, Pred_T 2.0e-5, Pred_s 2.1e-5.  Bars = 1.5x the flips and
    # 3x the output differences, so a regression of the fp32 path shows up here.
    assert flips / rows < 0.09
    bars = {"p_green_R": 8e-4, "p_red_R": 1.3e-3, "f_green_R": 1.5e-5, "f_red_R": 1.5e-5, "Pred_T": 6e-5, "Pred_s": 6.5e-5}
    for n in NAMES:
        assert diffs[n] < bars[n], (n, diffs[n])


def _train_module(cuda):
    from hspose_b200.HSPose import HSPose
    F = _flags()
    F.train, F.gcn_n_num = 1, 20
    net = fill_params(HSPose("PoseNet_only")).to(cuda).train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return net


def test_hspose_train_step_teacher_forced(cuda, golden):
    """Train-mode forward (batch-stat BN) + native fs_net losses + backward; RF-F tables forced."""
    from hspose_b200 import gcn3d
    g = golden("e2e_train")
    F = _flags()
    saved = {n: getattr(F, n) for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro")}
    for n in saved:
        setattr(F, n, 0.0)
    try:
        net = _train_module(cuda)
        batch = {k: v.to(cuda) for k, v in synth_batch(4, 1028, seed=2, train=True).items()}
        rf = [torch.from_numpy(g[f"rf{i}"].astype(np.int64)) for i in range(4)]
        # the golden ran on CPU where the augmentation's torch.rand draws advance the generator
        # Pool_layer's randperm uses; replay them so both sides pool the same rows
        torch.manual_seed(4321)
        for shape in [(4, 1)] * 6 + [(4, 1028, 3)]:
            torch.rand(shape)
        with gcn3d.force_rf_indices(rf):
            out, losses = net(**batch, do_loss=True)
        for n in NAMES:
            np.testing.assert_allclose(out[n].detach().cpu().numpy(), g["out_" + n], atol=1e-4, err_msg=n)  # bn3 over 4 samples
        for n in ("recon", "face_dis", "face_f", "face_normal"):
            np.testing.assert_allclose(out[n][:, ::64].detach().cpu().numpy(), g["out_" + n],
                                       atol=1e-3 if n == "face_normal" else 2e-4, err_msg=n)
        total = 0
        for k, v in losses["fsnet_loss"].items():
            np.testing.assert_allclose(v.item(), g["loss_fs_" + k].item(), rtol=1e-4, atol=1e-5, err_msg=k)
            total = total + v.reshape(())
        total.backward()
        names = list(g["grad_norm_names"])
        vals = g["grad_norm_values"]
        params = dict(net.named_parameters())
        worst = 0.0
        for n, v in zip(names, vals):
            mine = params[str(n)].grad.norm().item()
            worst = max(worst, abs(mine - v) / max(v, 1e-6))
            # biases in front of a batch-stat BN have zero analytic gradient (rounding noise ~3e-5)
            assert abs(mine - v) <= 2e-3 * v + 1e-4, (n, mine, v)
        # Element-wise gradients.  Only the three pose heads feed this loss, and each routes its
        # gradient through max-over-points: 3 x 4 x 256 winning rows carry the whole backbone
        # gradient, so one ReLU-mask / argmax near-tie decided differently by fp32 rounding moves
        # every backbone gradient by ~0.1-1 % (measured on B200, tools/debug_grad.py: the
        # reference's own fp32 gradients are 2e-3..9e-3 away, in relative L2, from the same graph
        # evaluated in fp64; ours are 3e-4..3e-3 away).  The bar is therefore a relative-L2 bound
        # at that noise floor; test_train_gradients_vs_fp64_oracle holds the tight bound.
        for key in g:
            if key.startswith("grad::"):
                mine = params[key[len("grad::"):]].grad.double().cpu().numpy()
                ref = g[key].astype(np.float64)
                if np.linalg.norm(ref) < 1e-3:     # zero analytic gradient (bias in front of a BN)
                    assert np.abs(mine - ref).max() < 1e-3, key
                    continue
                rel = np.linalg.norm(mine - ref) / np.linalg.norm(ref)
                assert rel < 2e-2, (key, rel)
        sd = net.state_dict()
        for key in g:
            if key.startswith("post::"):
                np.testing.assert_allclose(sd[key[len("post::"):]].cpu().numpy(), g[key], atol=1e-5)
    finally:
        for n, v in saved.items():
            setattr(F, n, v)


def test_train_gradients_vs_fp64_oracle(cuda, golden):
    """Backward parity at full depth: every backbone gradient of the teacher-forced train step
    against the SAME graph evaluated by the materialising oracle in fp64 on the device (the
    closest thing to the exact gradient).  Ours must be at least as close to it as a plain
    fp32 evaluation of the reference algorithm (oracle in fp32) is, up to a factor 3 (at most two
    parameters may sit outside that because of an isolated near-tie), and within 2e-2 in
    relative L2 outright."""
    from hspose_b200 import gcn3d
    from hspose_b200.HSPose import control_loss
    from hspose_b200.losses import fs_net_loss, get_gt_v
    from oracle import torch_oracle as to
    g = golden("e2e_train")
    F = _flags()
    saved = {n: getattr(F, n) for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro")}
    for n in saved:
        setattr(F, n, 0.0)
    try:
        net = _train_module(cuda)
        batch = {k: v.to(cuda) for k, v in synth_batch(4, 1028, seed=2, train=True).items()}
        rf = [torch.from_numpy(g[f"rf{i}"].astype(np.int64)).to(cuda) for i in range(4)]
        torch.manual_seed(99)
        with gcn3d.force_rf_indices(rf):
            _, losses = net(**batch, do_loss=True)
        sum(v.reshape(()) for v in losses["fsnet_loss"].values()).backward()
        mine = {n: p.grad.double().cpu().numpy() for n, p in net.named_parameters() if p.grad is not None}
        torch.manual_seed(99)
        samples = (torch.randperm(1028)[:257].to(cuda), torch.randperm(257)[:64].to(cuda))

        def oracle(dtype):
            sd = {}
            for name, t in net.state_dict().items():
                t = t.detach().clone()
                if t.is_floating_point():
                    t = t.to(dtype)
                    if name.rsplit(".", 1)[-1] not in ("running_mean", "running_var"):
                        t.requires_grad_(True)
                sd[name] = t
            b = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}
            out = to.posenet9d(sd, b["PC"], b["obj_id"], k=20, S=7, train=True, bn_training=True,
                               samples=samples, rf_indices=rf)
            green, red = get_gt_v(b["gt_R"])
            pred = {"Rot1": out["p_green_R"], "Rot1_f": out["f_green_R"], "Rot2": out["p_red_R"],
                    "Rot2_f": out["f_red_R"], "Recon": out["recon"], "Tran": out["Pred_T"],
                    "Size": out["Pred_s"]}
            gt = {"Rot1": green, "Rot2": red, "Recon": b["PC"], "Tran": b["gt_t"], "Size": b["gt_s"]}
            ls = fs_net_loss()(control_loss("PoseNet_only")[0], pred, gt, b["sym"])
            sum(v.reshape(()) for v in ls.values()).backward()
            return {n: t.grad.double().cpu().numpy() for n, t in sd.items()
                    if t.is_floating_point() and t.grad is not None}

        g64, g32 = oracle(torch.float64), oracle(torch.float32)
        checked = loose = 0
        for n, ref in g64.items():
            if "face_recon.conv_" not in n and "face_recon.bn" not in n:
                continue
            nr = np.linalg.norm(ref)
            if nr < 1e-3:
                continue
            e_mine = np.linalg.norm(mine["posenet." + n[len("posenet."):]] - ref) / nr
            e_f32 = np.linalg.norm(g32[n] - ref) / nr
            assert e_mine < 2e-2, (n, e_mine, e_f32)
            loose += e_mine > max(3 * e_f32, 5e-3)   # a near-tie decided differently: isolated
            checked += 1
        assert checked >= 20 and loose <= 2, (checked, loose)
    finally:
        for n, v in saved.items():
            setattr(F, n, v)


def test_api_functions_match_oracle(cuda):
    """The gcn3d function surface (names / shapes / dtypes of reference gcn3d.py:15-59,189-218)."""
    from hspose_b200 import gcn3d
    from oracle import c_oracle as co
    g = torch.Generator().manual_seed(8)
    v = torch.randn(2, 300, 3, generator=g) * 0.05
    vc = v.to(cuda)
    idx = gcn3d.get_neighbor_index(vc, 12)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (2, 300, 12)
    assert np.array_equal(idx.cpu().numpy(), co.neighbor_index(v.numpy(), 12))
    src = vc[:, :75].contiguous()
    nn = gcn3d.get_nearest_index(vc, src)
    assert tuple(nn.shape) == (2, 300, 1)
    assert np.array_equal(nn.cpu().numpy(), co.nearest_index(v.numpy(), v[:, :75].numpy()))
    feat = torch.randn(2, 300, 64, generator=g)
    rows = gcn3d.indexing_neighbor_new(feat.to(cuda), idx)
    ref = np.stack([feat[b].numpy()[idx[b].cpu().numpy()] for b in range(2)])
    assert np.array_equal(rows.cpu().numpy(), ref)
    d, i2 = gcn3d.get_receptive_fields(12, vc, mode='RF-P')
    assert np.array_equal(i2.cpu().numpy(), idx.cpu().numpy())
    np.testing.assert_allclose(d.cpu().numpy(), co.direction_norm(v.numpy(), idx.cpu().numpy()), atol=1e-6)
    G = gcn3d.get_ORL_global(feat.to(cuda), vc, 12)
    assert tuple(G.shape) == (2, 300, 64)
    np.testing.assert_allclose(G[:, 0].cpu().numpy(), co.orl_global_fwd(feat.numpy(), idx.cpu().numpy()), atol=1e-5)


@pytest.mark.parametrize("groups,dense_bar", [(("fsnet",), 0.97), (("fsnet", "recon", "geo", "prop"), 0.80)])
def test_mixed_precision_step_close_to_fp32(cuda, groups, dense_bar):
    """bf16 fast path (autocast) vs the fp32 path of the same module on the same batch, with
    the fp32 run's RF-F neighbour tables forced into the bf16 run (removes the KNN
    discontinuity): losses within bf16 tolerance, gradients with high cosine similarity.
    With the recon_6face voting terms in the objective the dense-path gradient also carries the
    weighted plane fit (a 3x3 inverse of sums over 1028 points: ill-conditioned in fp32 already,
    tests/test_losses_cpu.py), so bf16 activations move it more (measured cosine 0.89)."""
    from hspose_b200 import gcn3d
    from hspose_b200.HSPose import HSPose
    rf = []
    F = _flags()
    saved = {n: getattr(F, n) for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro")}
    for n in saved:
        setattr(F, n, 0.0)
    try:
        res = {}
        for mode in ("fp32", "bf16"):
            F.train, F.gcn_n_num = 1, 20
            net = fill_params(HSPose("PoseNet_only", chamfer_w=1.0, loss_groups=groups)).to(cuda).train()
            for m in net.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            batch = {k: v.to(cuda) for k, v in synth_batch(8, 1028, seed=3, train=True).items()}
            torch.manual_seed(99)
            ctx = gcn3d.record_rf_indices(rf) if mode == "fp32" else gcn3d.force_rf_indices(rf)
            with ctx, torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                out, losses = net(**batch, do_loss=True)
            total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
            total.backward()
            named = {n: p.grad.reshape(-1).float() for n, p in net.named_parameters() if p.grad is not None}
            res[mode] = (total.item(), named, {k: v.item() for k, v in losses["fsnet_loss"].items()})
        t32, g32, l32 = res["fp32"]
        t16, g16, l16 = res["bf16"]
        assert abs(t32 - t16) / abs(t32) < 5e-2, (t32, t16, l32, l16)

        def cos(pred):
            a = torch.cat([v for n, v in g32.items() if pred(n)])
            b = torch.cat([g16[n] for n in g32 if pred(n)])
            return torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        # dense per-point path (backbone forward -> conv1d_block -> recon_head -> Chamfer/recon
        # terms): smooth in the activations, so bf16 must track fp32 closely
        dense = cos(lambda n: ("conv1d_block" in n or "recon_head" in n) and n.endswith("weight"))
        assert dense > dense_bar, dense
        # everything else is routed through max-over-points and a batch-stat BN over the 8 objects
        # of this batch (PoseR.py:29-35): discontinuous in the activations, bf16 rounding re-routes
        # winners (measured per-parameter cosines 0.5-0.98, tools/debug_mixed.py) — sanity bound only
        assert cos(lambda n: True) > 0.6
    finally:
        for n, v in saved.items():
            setattr(F, n, v)


def test_cuda_graph_step_matches_eager(cuda):
    """engine.TrainStep: six optimiser steps replayed from one CUDA graph follow the eagerly launched
    loop (same CPU-generator pooling permutations, dropout/augmentation off).  The graph warm-up is
    side-effect free, so the graph run starts from identical weights and optimiser state."""
    from hspose_b200.engine import TrainStep
    from hspose_b200.HSPose import HSPose
    F = _flags()
    saved = {n: getattr(F, n) for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro")}
    for n in saved:
        setattr(F, n, 0.0)

    def run(graph):
        F.train, F.gcn_n_num = 1, 20
        net = fill_params(HSPose("PoseNet_only", chamfer_w=1.0)).to(cuda).train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        tr = TrainStep(net, lr=1e-5, amp=True, graph=graph)
        torch.manual_seed(77)
        out = []
        for i in range(6):   # graph mode: the first call warms up (side-effect free) and captures
            out.append(tr(synth_batch(4, 1028, seed=10 + (i % 2), train=True)).item())
        if graph:
            assert tr.launches_per_step and tr.launches_per_step > 50
        return np.array(out)
    try:
        e, g = run(False), run(True)
        assert np.all(np.isfinite(e)) and np.all(np.isfinite(g))
        # The graph run starts from identical weights AND identical optimiser state (the warm-up is side-effect
        # free), and every kernel of the mixed-precision step adds in a fixed order: the replayed loop follows the
        # eagerly launched one bit for bit over all six optimiser steps.  (With the float-atomics backward of mid
        # round 2 the two drifted apart by 1e-2 from step 2 on.)
        # measured: bit-identical (np.array_equal) — the bar leaves one part in a million
        assert np.all(np.abs(e - g) <= 1e-6 * np.abs(e)), (e, g)
    finally:
        for n, v in saved.items():
            setattr(F, n, v)


def test_cuda_graph_replay_equals_eager_single_step(cuda):
    """Same weights, same batch, same pooling rows: one replayed step's loss == one eager step's."""
    from hspose_b200.engine import TrainStep
    from hspose_b200.HSPose import HSPose
    F = _flags()
    saved = {n: getattr(F, n) for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro")}
    for n in saved:
        setattr(F, n, 0.0)
    try:
        F.train, F.gcn_n_num = 1, 20
        net = fill_params(HSPose("PoseNet_only", chamfer_w=1.0)).to(cuda).train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        batch = synth_batch(4, 1028, seed=5, train=True)
        tr = TrainStep(net, lr=0.0, amp=True, graph=True)    # lr 0: weights never move
        torch.manual_seed(5)
        tr(batch)                                            # warm-up + capture + 1 replay
        torch.manual_seed(6)
        l_graph = tr(batch).item()
        tr.use_graph = False
        torch.manual_seed(6)
        l_eager = tr(batch).item()
        assert abs(l_graph - l_eager) <= 2e-3 * abs(l_eager), (l_graph, l_eager)
    finally:
        for n, v in saved.items():
            setattr(F, n, v)


def test_fused_loss_kernel_matches_loss_modules(cuda):
    """K8 (csrc/losses.cu: one templated per-object function evaluated with forward-mode dual numbers +
    two point passes) against the tensor-algebra loss modules of losses.py — which tests/test_losses_cpu.py
    pins to the reference's own modules — on synthetic predictions covering every symmetry branch:
    all 19 terms and the gradients w.r.t. every prediction."""
    import hspose_b200.ops as ops
    from hspose_b200.losses import fs_net_loss, geo_transform_loss, prop_rot_loss, recon_6face_loss
    from hspose_b200.geom import get_gt_v
    from hspose_b200.synth import synth_predictions
    F = _flags()
    B, N = 12, 257
    pred, gt = synth_predictions(B, N, seed=11)
    g = torch.Generator().manual_seed(5)
    face_raw = torch.cat([(pred["face_normal"] * (0.5 + torch.rand(B, N, 6, 1, generator=g))).reshape(B, N, 18),
                          pred["face_dis"], torch.logit(pred["face_f"])], dim=2)
    dev = {k: v.to(cuda) for k, v in gt.items()}

    def leaves():
        out = {k: pred[k].to(cuda).clone().requires_grad_() for k in
               ("recon", "p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s")}
        out["face"] = face_raw.to(cuda).clone().requires_grad_()
        return out
    # reference path: the modules, from the same raw face tensor (normalise / sigmoid as PoseNet9D.py:28-33)
    a = leaves()
    fn = a["face"][:, :, :18].view(B, N, 6, 3)
    fn = fn / torch.norm(fn, dim=-1, keepdim=True)
    fd, fc = a["face"][:, :, 18:24], torch.sigmoid(a["face"][:, :, 24:])
    gg, gr = get_gt_v(dev["gt_R"])
    names = (['Rot1', 'Rot2', 'Rot1_cos', 'Rot2_cos', 'Rot_regular', 'Tran', 'Size', 'R_con'],
             ['Per_point', 'Point_voting'], ['Geo_point'], ['Prop_pm', 'Prop_sym'])
    ref = {}
    ref.update(fs_net_loss()(names[0], {'Rot1': a["p_green_R"], 'Rot1_f': a["f_green_R"], 'Rot2': a["p_red_R"],
                                        'Rot2_f': a["f_red_R"], 'Recon': a["recon"], 'Tran': a["Pred_T"],
                                        'Size': a["Pred_s"]},
                             {'Rot1': gg, 'Rot2': gr, 'Recon': dev["PC"], 'Tran': dev["gt_t"], 'Size': dev["gt_s"]},
                             dev["sym"]))
    ref.update(recon_6face_loss().to(cuda)(names[1], {'F_n': fn, 'F_d': fd, 'F_c': fc, 'Rot1': a["p_green_R"],
                                                      'Rot1_f': a["f_green_R"].detach(), 'Rot2': a["p_red_R"],
                                                      'Rot2_f': a["f_red_R"].detach(), 'Tran': a["Pred_T"],
                                                      'Size': a["Pred_s"]},
                                           {'R': dev["gt_R"], 'T': dev["gt_t"], 'Size': dev["gt_s"],
                                            'Mean_shape': dev["mean_shape"], 'Points': dev["PC"]}, dev["sym"],
                                           dev["obj_id"]))
    ref.update(geo_transform_loss()(names[2], {'Rot1': a["p_green_R"], 'Rot2': a["p_red_R"], 'Tran': a["Pred_T"]},
                                    {'Points': dev["PC"], 'R': dev["gt_R"], 'T': dev["gt_t"]}, dev["sym"]))
    ref.update(prop_rot_loss().to(cuda)(names[3], {'Recon': a["recon"], 'Rot1': a["p_green_R"], 'Rot2': a["p_red_R"],
                                                   'Tran': a["Pred_T"], 'Rot1_f': a["f_green_R"].detach(),
                                                   'Rot2_f': a["f_red_R"].detach()},
                                        {'Points': dev["PC"], 'R': dev["gt_R"], 'T': dev["gt_t"]}, dev["sym"]))
    b = leaves()
    w = [getattr(F, n) for n in ops.LOSS_WEIGHT_FLAGS]
    got = ops.fused_losses(w, b["face"], b["recon"], b["p_green_R"], b["p_red_R"], b["f_green_R"], b["f_red_R"],
                           b["Pred_T"], b["Pred_s"], dev["PC"], dev["gt_R"], dev["gt_t"], dev["gt_s"],
                           dev["mean_shape"], dev["sym"], dev["obj_id"])
    assert set(got) == set(ref), (sorted(got), sorted(ref))
    for k in got:
        r = ref[k].item()
        assert abs(got[k].item() - r) <= 2e-5 * max(1.0, abs(r)), (k, got[k].item(), r)
    # gradients, term group by term group (so a wrong term cannot hide behind a large one)
    groups = {"fs": names[0], "recon": ["recon_"], "geo": ["geo_"], "prop": ["Prop_"]}
    for gname in ("fs", "recon", "geo", "prop"):
        keys = [k for k in got if (k.startswith(tuple(groups[gname])) if gname != "fs" else
                                   k in ("Rot1", "Rot1_cos", "Rot2", "Rot2_cos", "Rot_r_a", "Tran", "Size", "R_con"))]
        for d in (a, b):
            for v in d.values():
                v.grad = None
        sum(ref[k] for k in keys).backward(retain_graph=True)
        sum(got[k] for k in keys).backward(retain_graph=True)
        for name in a:
            ga, gb = a[name].grad, b[name].grad
            if ga is None or float(ga.abs().max()) == 0.0:
                assert gb is None or float(gb.abs().max()) == 0.0, (gname, name)
                continue
            assert gb is not None, (gname, name)
            # plane-fit gradients: fp32 noise floor 1e-4..3e-4 of the tensor max (tests/test_losses_cpu.py)
            rel = 1e-3 if gname == "recon" and name in ("face", "Pred_T") else 5e-5
            err = (ga - gb).abs().max().item()
            assert err <= rel * ga.abs().max().item() + 1e-9, (gname, name, err, ga.abs().max().item())


@pytest.mark.parametrize("groups", [("fsnet", "recon", "geo", "prop"), ("fsnet", "geo")])
def test_loss_total_is_the_sum_of_the_returned_terms(cuda, groups):
    """HSPose's fused loss path also returns `.total` (one masked reduction of the term vector + Chamfer) for train
    loops that only need the sum (engine.TrainStep): same value and same gradients as adding the dict entries one by
    one the way engine/train.py:137-149 does; the 4-key dict itself is unchanged."""
    from hspose_b200.HSPose import HSPose
    F = _flags()
    F.train, F.gcn_n_num = 1, 20
    net = fill_params(HSPose("PoseNet_only", chamfer_w=1.0, loss_groups=groups)).to(cuda).train()
    batch = {k: v.to(cuda) for k, v in synth_batch(4, 1028, seed=5, train=True).items()}
    grads = []
    for use_total in (True, False):
        torch.manual_seed(7)
        net.zero_grad(set_to_none=True)
        _, losses = net(**batch, do_loss=True)
        assert set(losses) == {"fsnet_loss", "recon_loss", "geo_loss", "prop_loss"}
        manual = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
        assert losses.total is not None
        assert abs(float(losses.total) - float(manual)) <= 1e-5 * abs(float(manual))
        (losses.total if use_total else manual).backward()
        grads.append(torch.cat([p.grad.flatten() for p in net.parameters() if p.grad is not None]))
    rel = float((grads[0] - grads[1]).norm() / grads[1].norm())
    assert rel <= 1e-4, rel        # fp32 float-atomic noise between two runs of the same step is ~1e-6


@pytest.mark.parametrize("tag,probs", [
    ("all", dict(aug_pc_pro=1.0, aug_rt_pro=1.0, aug_bb_pro=1.0, aug_bc_pro=1.0)),
    ("default", dict(aug_pc_pro=0.2, aug_rt_pro=0.3, aug_bb_pro=0.3, aug_bc_pro=0.3))])
def test_augment_kernel_matches_reference(cuda, golden, tag, probs):
    """K10 (csrc/augment.cu) against goldens of the REFERENCE's HSPose.data_augment (network/HSPose.py:185-256),
    fed the same uniform draws the reference consumed (CPU generator, seed 777, same order and shapes)."""
    import hspose_b200.ops as ops
    g = golden("aug")
    b = synth_batch(8, 1028, seed=5, train=True)
    torch.manual_seed(777)
    bs = 8
    g_bb, g_rt, g_bc = (torch.rand((bs, 1)) for _ in range(3))
    ey_up, ey_down = torch.rand((bs, 1)), torch.rand((bs, 1))
    g_pc = torch.rand((bs, 1))
    defor = torch.rand(b["PC"].shape)
    d = {k: v.to(cuda) for k, v in b.items()}
    PC, R, t, s = ops.augment(d["PC"], d["gt_R"], d["gt_t"], d["gt_s"], d["mean_shape"], d["sym"], d["aug_bb"],
                              d["aug_rt_t"], d["aug_rt_r"], d["model_point"], d["nocs_scale"], d["obj_id"],
                              torch.cat([g_bb, g_rt, g_bc, g_pc], 1).to(cuda), torch.cat([ey_up, ey_down], 1).to(cuda),
                              defor.to(cuda), (probs["aug_bb_pro"], probs["aug_rt_pro"], probs["aug_bc_pro"],
                                               probs["aug_pc_pro"]), 0.2)
    np.testing.assert_allclose(PC.cpu().numpy(), g[f"{tag}_PC"], atol=2e-6)
    np.testing.assert_allclose(R.cpu().numpy(), g[f"{tag}_R"], atol=1e-6)
    np.testing.assert_allclose(t.cpu().numpy(), g[f"{tag}_t"], atol=1e-6)
    np.testing.assert_allclose(s.cpu().numpy(), g[f"{tag}_s"], atol=1e-6)


def test_eval_runner_buckets_and_generate_RT(cuda, golden):
    """engine.EvalRunner (the per-image call of evaluation/evaluate.py:91-108 as one CUDA-graph replay per
    batch-size bucket): outputs equal the eager forward for batch sizes inside and between buckets, pred_RT equals
    geom.generate_RT of the same outputs, a padded bucket does not disturb the real objects."""
    from hspose_b200 import geom
    from hspose_b200.engine import EvalRunner
    from hspose_b200.HSPose import HSPose
    F = _flags()
    F.train, F.gcn_n_num = 0, 20
    try:
        model = HSPose("PoseNet_only")
        model.posenet = _posenet(cuda, 0, 20)
        model = model.to(cuda).eval()
        runner = EvalRunner(model, buckets=(1, 4, 8))
        for B in (1, 3, 4, 6):
            b = synth_batch(B, 1028, seed=20 + B, train=False)
            args = [b[k].to(cuda) for k in ("PC", "obj_id", "mean_shape", "sym")]
            runner(*args)                                     # first use of a bucket: warm-up + capture (extra draws)
            torch.manual_seed(7)
            got = runner(*args)
            torch.manual_seed(7)
            got2 = runner(*args)                              # replay of the captured graph, same permutations
            nb = next(x for x in (1, 4, 8) if x >= B)
            padded = [torch.cat([a, a[:1].expand(nb - B, *a.shape[1:])]) if nb > B else a for a in args]
            torch.manual_seed(7)
            with torch.no_grad():
                ref = model(PC=padded[0], obj_id=padded[1], mean_shape=padded[2], sym=padded[3])
                torch.manual_seed(7)
                loose = model(PC=args[0], obj_id=args[1], mean_shape=args[2], sym=args[3])
            for n in NAMES:
                assert got[n].shape == loose[n].shape
                # the replayed graph == the eager forward of the same (padded) batch
                assert (got[n] - ref[n][:B]).abs().max().item() <= 1e-6, (B, n)
                assert torch.equal(got[n], got2[n]), (B, n)
                # vs the UNPADDED eager batch only the library GEMMs' blocking changes (other row count): the
                # feature-space KNN re-decides a few neighbour sets, as the reference itself does when its batch
                # size changes (6e-4 .. 2e-3 on the rotation vectors)
                assert (got[n] - loose[n]).abs().max().item() <= 1e-2, (B, n)
            RT = geom.generate_RT([got["p_green_R"], got["p_red_R"]], [got["f_green_R"], got["f_red_R"]],
                                  got["Pred_T"], "vec", args[3])
            assert torch.allclose(got["pred_RT"], RT, atol=1e-6)
            assert torch.allclose(got["pred_s"], got["Pred_s"] + args[2])
        assert sorted(runner.graphs) == [1, 4, 8]
    finally:
        F.train, F.gcn_n_num = 1, 20


def test_amp_train_gradients_are_bit_reproducible(cuda):
    """SURVEY.md §5 (sort/segment based rather than float atomics): on the mixed-precision train path every backward
    kernel adds in a fixed order (K4b fixed-point slabs, gathered pool / up-sample / Chamfer backward, fixed-order
    reductions), so two runs of the same forward + backward give bit-identical gradients for ALL parameters."""
    from hspose_b200 import parallel
    from hspose_b200.HSPose import HSPose
    F = _flags()
    saved = {n: getattr(F, n) for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro", "train", "gcn_n_num")}
    for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro"):
        setattr(F, n, 0.0)
    F.train, F.gcn_n_num = 1, 20
    try:
        model = fill_params(HSPose("PoseNet_only", chamfer_w=1.0)).to(cuda).train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        batch = {k: v.to(cuda) for k, v in synth_batch(6, 1028, seed=9, train=True).items()}
        runs = []
        for r in range(3):
            parallel.seed_all(4321)
            for p in model.parameters():
                p.grad = None
            junk = torch.randn(1 << (20 + r), device=cuda)      # shift the allocator between the runs
            with torch.autocast("cuda", dtype=torch.bfloat16):
                _, losses = model(**batch, do_loss=True)
            total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
            total.backward()
            del junk
            runs.append({n: p.grad.detach().clone() for n, p in model.posenet.named_parameters() if p.grad is not None})
        assert len(runs[0]) > 100
        bad = [n for n in runs[0] if not (torch.equal(runs[0][n], runs[1][n]) and torch.equal(runs[0][n], runs[2][n]))]
        assert not bad, bad
    finally:
        for n, v in saved.items():
            setattr(F, n, v)
