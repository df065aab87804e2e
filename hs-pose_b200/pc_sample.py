"""Input pre-stage on the device (K11): depth ROI -> camera-frame cloud -> `random_points` sampled points.

Mirrors the two places the reference builds the network's input cloud:

* `PC_sample(obj_mask, Depth, camK, coor2d)` — network/point_sample/pc_sample.py:8-77, the `depth=` path of
  `HSPose.forward` (network/HSPose.py:40-48).  Same name, arguments and return value (`PC / 1000.0`, (B,n,3)).
  The reference loops over the batch in Python (boolean-mask indexing = one host sync per object); here one
  launch compacts the whole batch, ONE count read-back sizes the draws, and the indices are drawn on the host
  with `np.random.choice` in the reference's order — so with the same numpy seed the result is the reference's.
* `depth_to_pcl` + `sample_points` — `PoseDataset._depth_to_pcl` / `_sample_points` (datasets/load_data.py:322-333,
  307-320; numpy float64 arithmetic, `/ 1000.0` from :277), batched.  `sample_points(..., ids=None)` uses the
  device rule (no host round trip, CUDA-graph safe): tile when short, random subset when long.

No CPU path: inputs must be CUDA tensors (ops._need raises otherwise).
"""
import numpy as np
import torch

from . import ops
from .flags import FLAGS


def PC_sample(obj_mask, Depth, camK, coor2d):
    """:param obj_mask: (B,1,H,W) / (B,H,W) mask, or (B,2,H,W) predicted logits (arg-max channel 1 = object)
    :param Depth: (B,1,H,W) millimetres; camK (B,3,3); coor2d (B,2,H,W)
    :return: (B, FLAGS.random_points, 3) metres, or None when an object has <= 1 valid pixel (the reference
             returns `(None, None)` there, which its caller cannot handle — network/HSPose.py:46-48 tests `is None`)."""
    if obj_mask.dim() == 4 and obj_mask.shape[1] == 2:          # predicted mask (pc_sample.py:16-18)
        obj_mask = torch.max(torch.softmax(obj_mask, dim=1), dim=1)[1]
    if FLAGS.sample_method != 'basic':
        raise NotImplementedError
    n = int(FLAGS.random_points)
    cloud, count = ops.depth_to_cloud(Depth, obj_mask, coor2d, camK.float())
    totals = count.cpu().tolist()                                # the one host sync
    choose = np.empty((len(totals), n), dtype=np.int32)
    for i, l_all in enumerate(totals):
        if l_all <= 1.0:
            return None
        choose[i] = np.random.choice(l_all, n, replace=l_all < n)   # pc_sample.py:61-66, same RNG stream
    return ops.sample_points(cloud, count, n, choose=torch.from_numpy(choose).to(cloud.device, non_blocking=True))


def depth_to_pcl(depth, K, xymap, mask):
    """Batched `PoseDataset._depth_to_pcl(...) / 1000.0`: -> (cloud (B,H*W,3), count (B,)).  K is used in float64
    like numpy does (pass the float64 intrinsics for bit-exact parity)."""
    return ops.depth_to_cloud(depth, mask, xymap, K.double())


def sample_points(cloud, count, n_pts, ids=None, seed=0):
    """Batched `PoseDataset._sample_points`.  ids (B,n_pts): the host's `np.random.permutation(total)[:n_pts]`
    (or arange/tile indices) for parity runs; None: the device rule."""
    return ops.sample_points(cloud, count, n_pts, choose=ids, seed=seed)
