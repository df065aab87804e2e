"""Seeded synthetic inputs and deterministic parameter fill (no oracle arithmetic here).

Shared by tests/golden/make_golden.py (which runs the real reference), the
parity tests, bench.py and the profiling tools so that every side sees byte-identical
inputs and weights without shipping multi-megabyte checkpoints or datasets.  Input statistics follow
SURVEY.md §8(d) (object of ~10-30 cm at 0.8 m; reference datasets/load_data.py).
"""
import math
import zlib

import torch

# per-category symmetry flags and mean shapes (metres), restated from the
# reference's dataset tables (datasets/load_data.py:358-381, :421-436):
# bottle, bowl, camera, can, laptop, mug
_SYM = torch.tensor([[1, 1, 0, 1], [1, 1, 0, 1], [0, 0, 0, 0],
                     [1, 1, 1, 1], [0, 1, 0, 0], [0, 1, 0, 0]], dtype=torch.float32)
_MEAN_SHAPE = torch.tensor([[87, 220, 89], [165, 80, 165], [88, 128, 156],
                            [68, 146, 72], [346, 200, 335], [146, 83, 114]],
                           dtype=torch.float32) / 1000.0


def random_rotations(n, gen, max_deg=None):
    if max_deg is None:
        q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=gen))
        q = q * torch.sign(torch.diagonal(r, dim1=1, dim2=2)).unsqueeze(1)
        det = torch.linalg.det(q)
        q[:, :, 2] *= det.unsqueeze(-1)
        return q.contiguous()
    a = (torch.rand(n, 3, generator=gen) * 2 - 1) * math.radians(max_deg)
    cx, cy, cz = torch.cos(a).unbind(1)
    sx, sy, sz = torch.sin(a).unbind(1)
    one, zero = torch.ones(n), torch.zeros(n)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).view(n, 3, 3)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).view(n, 3, 3)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).view(n, 3, 3)
    return (Rz @ Ry @ Rx).contiguous()


def synth_batch(B, N=1028, seed=1, train=True):
    """Dict of fp32 CPU tensors with the kwargs HSPose.forward takes."""
    g = torch.Generator().manual_seed(seed)
    PC = torch.randn(B, N, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])
    cat = torch.randint(0, 6, (B,), generator=g)
    out = {
        "PC": PC.contiguous(),
        "obj_id": cat.float() if train else cat,
        "sym": _SYM[cat].clone(),
        "mean_shape": _MEAN_SHAPE[cat].clone(),
        "gt_R": random_rotations(B, g),
        "gt_t": torch.tensor([0.0, 0.0, 0.8]).repeat(B, 1),
        "gt_s": torch.randn(B, 3, generator=g) * 0.01,
        "aug_bb": torch.rand(B, 3, generator=g) * 0.4 + 0.8,
        "aug_rt_t": torch.rand(B, 3, generator=g) * 0.1 - 0.05,
        "aug_rt_r": random_rotations(B, g, max_deg=15.0),
        "model_point": torch.randn(B, 1024, 3, generator=g) * 0.2,
        "nocs_scale": torch.rand(B, generator=g) * 0.3 + 0.1,
    }
    return out


def synth_predictions(B, N, seed):
    """Random but plausible network outputs + ground truth for the loss goldens (all six
    categories present so every symmetry branch of the losses runs)."""
    g = torch.Generator().manual_seed(seed)
    cat = torch.arange(B) % 6
    R = random_rotations(B, g)
    t = torch.tensor([0.0, 0.0, 0.8]).repeat(B, 1) + torch.randn(B, 3, generator=g) * 0.02
    s = torch.randn(B, 3, generator=g) * 0.01
    ms = _MEAN_SHAPE[cat].clone()
    obj = (torch.rand(B, N, 3, generator=g) - 0.5) * (ms + s).unsqueeze(1)        # points inside the box
    PC = obj @ R.transpose(1, 2) + t.unsqueeze(1)
    nrm = lambda v: v / v.norm(dim=-1, keepdim=True)
    gt_axes = R.transpose(1, 2)                                                    # (B,3,3): axes[f] = R[:, f]
    face_axes = torch.cat([gt_axes[:, [1, 0, 2]], -gt_axes[:, [1, 2, 0]]], dim=1)  # network order y+,x+,z+,y-,z-,x-
    half = ((ms + s) / 2.0)
    d_plus, d_minus = half.unsqueeze(1) - obj, half.unsqueeze(1) + obj             # (B,N,3) for x,y,z
    face_d = torch.stack([d_plus[..., 1], d_plus[..., 0], d_plus[..., 2], d_minus[..., 1], d_minus[..., 2],
                          d_minus[..., 0]], dim=-1)
    pred = {
        "face_normal": nrm(face_axes.unsqueeze(1) + 0.15 * torch.randn(B, N, 6, 3, generator=g)),
        "face_dis": face_d + 0.01 * torch.randn(B, N, 6, generator=g),
        "face_f": torch.rand(B, N, 6, generator=g) * 0.9 + 0.05,
        "recon": PC + 0.01 * torch.randn(B, N, 3, generator=g),
        "p_green_R": nrm(R[:, :, 1] + 0.2 * torch.randn(B, 3, generator=g)),
        "p_red_R": nrm(R[:, :, 0] + 0.2 * torch.randn(B, 3, generator=g)),
        "f_green_R": torch.rand(B, generator=g) * 0.8 + 0.1,
        "f_red_R": torch.rand(B, generator=g) * 0.8 + 0.1,
        "Pred_T": t + 0.01 * torch.randn(B, 3, generator=g),
        "Pred_s": s + 0.005 * torch.randn(B, 3, generator=g),
    }
    gt = {"PC": PC, "gt_R": R, "gt_t": t, "gt_s": s, "mean_shape": ms, "sym": _SYM[cat].clone(),
          "obj_id": cat.float()}
    return pred, gt


def tiled_cloud(B, N=1028, unique=400, seed=7):
    """Stress set: `unique` points tiled to N (exact duplicates / distance ties)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, unique, 3, generator=g) * 0.05
    reps = (N + unique - 1) // unique
    return base.repeat(1, reps, 1)[:, :N].contiguous()


def fill_params(module, seed=0):
    """Overwrite every parameter/buffer of `module` with values that depend only
    on (key name, shape, seed) — identical for the reference module and ours as
    long as the state_dict keys/shapes agree (which is itself the drop-in
    contract)."""
    sd = module.state_dict()
    with torch.no_grad():
        for name in sorted(sd.keys()):
            t = sd[name]
            if not t.is_floating_point():
                continue  # num_batches_tracked
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
            u = torch.rand(t.shape, generator=g) * 2 - 1
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "running_var":
                v = 1.0 + 0.2 * u.abs()
            elif leaf == "running_mean":
                v = 0.1 * u
            elif t.dim() == 1 and leaf == "weight":  # BN affine scale
                v = 1.0 + 0.1 * u
            elif t.dim() == 1:  # biases
                v = 0.05 * u
            elif leaf in ("weights", "directions"):  # (Cin,(S+1)Cout) / (3,S*C)
                v = u / math.sqrt(t.shape[1])
            else:  # conv weights (out, in, 1)
                v = u / math.sqrt(t.shape[1])
            t.copy_(v.to(t.dtype))
    return module
