"""numpy restatement of the reference's input pre-stage — TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's
baseline legs may import it; the product path is hs-pose_b200/pc_sample.py over the K11 kernels).

Pinned by tests/golden/prestage.npz, generated from the reference's own functions
(`PoseDataset._depth_to_pcl`, `PoseDataset._sample_points`, `PC_sample`) by tests/golden/make_golden.py.
"""
import numpy as np


def depth_to_pcl(depth, K, xymap, mask):
    """datasets/load_data.py:322-333 (one object; float64 arithmetic, float32 result) followed by `/ 1000.0`
    of :277 (float32 / python float -> float32)."""
    K = np.asarray(K, dtype=np.float64).reshape(-1)
    cx, cy, fx, fy = K[2], K[5], K[0], K[4]
    d = depth.reshape(-1).astype(np.float64)
    valid = ((d > 0) * mask.reshape(-1)) > 0
    d = d[valid]
    x = (xymap[0].reshape(-1)[valid] - cx) * d / fx
    y = (xymap[1].reshape(-1)[valid] - cy) * d / fy
    return np.stack((x, y, d), axis=-1).astype(np.float32) / np.float32(1000.0)


def sample_points(pcl, n_pts, ids=None):
    """datasets/load_data.py:307-320.  ids = the permutation prefix the caller drew (None only when no draw
    is needed, i.e. total <= n_pts)."""
    total = pcl.shape[0]
    if total < n_pts:
        return np.concatenate([np.tile(pcl, (n_pts // total, 1)), pcl[:n_pts % total]], axis=0)
    if total > n_pts:
        return pcl[ids]
    return pcl


def pc_sample(obj_mask, depth, camK, coor2d, choose):
    """network/point_sample/pc_sample.py:24-77 for one object, float32 arithmetic; `choose` = the
    np.random.choice draw."""
    d = depth.astype(np.float32)
    fuse = (obj_mask.astype(np.float32) * (d > 0).astype(np.float32)) > 0
    K = camK.astype(np.float32)
    x = (coor2d[0].astype(np.float32) - K[0, 2]) * d / K[0, 0]
    y = (coor2d[1].astype(np.float32) - K[1, 2]) * d / K[1, 1]
    p = np.stack([x[fuse], y[fuse], d[fuse]], axis=1).astype(np.float32)
    return p[choose] / np.float32(1000.0)
