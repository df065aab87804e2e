"""CPU: the torch restatement (oracle/torch_oracle.py) and the native loss /
augmentation host code against end-to-end golden vectors produced by the REAL
reference (tests/golden/make_golden.py: e2e_eval, e2e_train)."""
import numpy as np
import pytest
import torch

from oracle import torch_oracle as to
from oracle.synth import fill_params, synth_batch

NAMES = ["p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s"]


def _module(train):
    """Our HSPose module is pure nn.Module state on CPU (only forward needs the GPU)."""
    import hspose_b200.flags as hf
    from hspose_b200.HSPose import HSPose
    hf.get_flags().train = train
    m = HSPose("PoseNet_only")
    hf.get_flags().train = 1
    return fill_params(m)


def _posenet(train):
    """PoseNet9D filled exactly as make_golden.golden_e2e_eval fills the reference's."""
    import hspose_b200.flags as hf
    from hspose_b200.PoseNet9D import PoseNet9D
    hf.get_flags().train = train
    m = PoseNet9D()
    hf.get_flags().train = 1
    return fill_params(m)


def test_state_dict_layout_matches_reference(golden):
    g = golden("e2e_eval")
    assert len(_posenet(0).state_dict()) == int(g["n_state_keys"]) == 107  # eval build
    sd = _module(1).state_dict()
    assert len(sd) == 160
    assert sum(p.numel() for p in _module(1).parameters()) == 9709871
    assert tuple(sd["posenet.face_recon.conv_1.weights"].shape) == (128, 1024)
    assert tuple(sd["posenet.face_recon.conv_1.directions"].shape) == (3, 896)
    assert tuple(sd["posenet.face_recon.conv_1.STE_layer.weight"].shape) == (128, 128, 1)
    assert tuple(sd["posenet.face_recon.conv_1.conv2.weight"].shape) == (128, 256, 1)


@pytest.mark.parametrize("k", [20, 16])
def test_oracle_eval_forward_teacher_forced(golden, k):
    """T2: RF-F tables from the reference run -> pose/size within 1e-5."""
    g = golden("e2e_eval")
    sd = _posenet(0).state_dict()
    batch = synth_batch(2, 1028, seed=1, train=False)
    rf = [torch.from_numpy(g[f"k{k}_rf{i}"].astype(np.int64)) for i in range(4)]
    samples = (torch.from_numpy(g[f"k{k}_sample1"].astype(np.int64)),
               torch.from_numpy(g[f"k{k}_sample2"].astype(np.int64)))
    with torch.no_grad():
        out = to.posenet9d(sd, batch["PC"], batch["obj_id"], k=k, train=False, samples=samples,
                           rf_indices=rf, pre="")
    for n in NAMES:
        np.testing.assert_allclose(out[n].numpy(), g[f"k{k}_{n}"], atol=1e-5, err_msg=n)
    np.testing.assert_allclose(out["feat"][:, ::16].numpy(), g[f"k{k}_feat_s16"], atol=2e-5)


def test_oracle_eval_forward_free_running(golden):
    """T3: free-running; the reference's own noise floor is ~1e-3 on rotations
    (SURVEY.md App. C.2), so only a loose bound is asserted and flips are counted."""
    g = golden("e2e_eval")
    sd = _posenet(0).state_dict()
    batch = synth_batch(2, 1028, seed=1, train=False)
    torch.manual_seed(1234)
    with torch.no_grad():
        out = to.posenet9d(sd, batch["PC"], batch["obj_id"], k=20, train=False, pre="")
    assert np.array_equal(out["samples"][0].numpy(), g["k20_sample1"])
    assert np.array_equal(out["samples"][1].numpy(), g["k20_sample2"])
    for n in NAMES:
        np.testing.assert_allclose(out[n].numpy(), g[f"k20_{n}"], atol=5e-3, err_msg=n)
    same = np.mean([(np.sort(out["rf_indices"][i].numpy(), -1) == np.sort(g[f"k20_rf{i}"], -1)).all(-1).mean()
                    for i in range(4)])
    assert same > 0.95


def test_native_losses_match_reference(golden):
    """hs-pose_b200/losses.py::fs_net_loss (pure torch) on the reference's own outputs."""
    g = golden("e2e_train")
    from hspose_b200.losses import fs_net_loss, get_gt_v
    from hspose_b200.HSPose import control_loss
    batch = synth_batch(4, 1028, seed=2, train=True)
    pred = {"Rot1": g["out_p_green_R"], "Rot1_f": g["out_f_green_R"], "Rot2": g["out_p_red_R"],
            "Rot2_f": g["out_f_red_R"], "Tran": g["out_Pred_T"], "Size": g["out_Pred_s"], "Recon": None}
    pred = {k: (torch.from_numpy(v) if v is not None else None) for k, v in pred.items()}
    green, red = get_gt_v(batch["gt_R"])
    gt = {"Rot1": green, "Rot2": red, "Recon": batch["PC"], "Tran": batch["gt_t"], "Size": batch["gt_s"]}
    out = fs_net_loss()(control_loss("PoseNet_only")[0], pred, gt, batch["sym"])
    keys = [k[len("loss_fs_"):] for k in g if k.startswith("loss_fs_")]
    assert sorted(keys) == sorted(out.keys())
    for k in keys:
        np.testing.assert_allclose(out[k].item(), g["loss_fs_" + k].item(), rtol=1e-5, atol=1e-6, err_msg=k)


def replay_train_samples(bs=4, n=1028, seed=4321):
    """The golden train step ran on CPU, where HSPose.data_augment's torch.rand calls
    (reference HSPose.py:233-246, data_augmentation.py:109-110,136) advance the SAME generator
    Pool_layer's randperm uses.  Replay those draws to recover the two pooling samples."""
    torch.manual_seed(seed)
    for shape in [(bs, 1)] * 6 + [(bs, n, 3)]:
        torch.rand(shape)
    s1 = torch.randperm(n)[: n // 4]
    s2 = torch.randperm(n // 4)[: (n // 4) // 4]
    return s1, s2


def test_oracle_train_forward_teacher_forced(golden):
    """Train-mode (batch-stat BN, dropout off) oracle forward vs the reference's HSPose run."""
    g = golden("e2e_train")
    sd = _module(1).state_dict()
    batch = synth_batch(4, 1028, seed=2, train=True)
    rf = [torch.from_numpy(g[f"rf{i}"].astype(np.int64)) for i in range(4)]
    with torch.no_grad():
        out = to.posenet9d({k: v.clone() for k, v in sd.items()}, batch["PC"], batch["obj_id"], k=20,
                           train=True, bn_training=True, samples=replay_train_samples(), rf_indices=rf)
    for n in NAMES:
        np.testing.assert_allclose(out[n].numpy(), g["out_" + n], atol=1e-4, err_msg=n)  # bn3 normalises over 4 samples
    for n in ("recon", "face_dis", "face_f", "face_normal"):
        np.testing.assert_allclose(out[n][:, ::64].numpy(), g["out_" + n],
                                   atol=1e-3 if n == "face_normal" else 1e-4, err_msg=n)  # fn/|fn|, |fn| small


@pytest.mark.parametrize("tag,probs", [("all", dict(aug_pc_pro=1.0, aug_rt_pro=1.0, aug_bb_pro=1.0, aug_bc_pro=1.0)),
                                       ("default", dict(aug_pc_pro=0.2, aug_rt_pro=0.3, aug_bb_pro=0.3, aug_bc_pro=0.3))])
def test_data_augment_matches_reference(golden, tag, probs):
    """HSPose.data_augment (hs-pose_b200/augment.py, row-vector algebra) against the reference's
    network/HSPose.py:185-256 run on CPU with the same seed: same RNG draws in the same order, same
    Bernoulli gating, same deformations (bounding-box scaling, rotation/translation, the mug/bowl taper,
    per-point noise) — with every branch forced on and with the shipped probabilities."""
    import hspose_b200.flags as hf
    from hspose_b200.HSPose import HSPose
    from hspose_b200.synth import synth_batch
    g = golden("aug")
    F = hf.get_flags()
    saved = {n: getattr(F, n) for n in probs}
    saved_train = F.train
    try:
        F.train = 1
        for n, v in probs.items():
            setattr(F, n, v)
        net = HSPose("PoseNet_only")
        b = synth_batch(8, 1028, seed=5, train=True)
        torch.manual_seed(777)
        with torch.no_grad():
            PC, R, t, s = net.data_augment(b["PC"].clone(), b["gt_R"].clone(), b["gt_t"].clone(), b["gt_s"].clone(),
                                           b["mean_shape"], b["sym"], b["aug_bb"], b["aug_rt_t"], b["aug_rt_r"],
                                           b["model_point"].clone(), b["nocs_scale"], b["obj_id"])
        np.testing.assert_allclose(PC.numpy(), g[f"{tag}_PC"], atol=2e-6)
        np.testing.assert_allclose(R.numpy(), g[f"{tag}_R"], atol=1e-6)
        np.testing.assert_allclose(t.numpy(), g[f"{tag}_t"], atol=1e-6)
        np.testing.assert_allclose(s.numpy(), g[f"{tag}_s"], atol=1e-6)
    finally:
        F.train = saved_train
        for n, v in saved.items():
            setattr(F, n, v)
