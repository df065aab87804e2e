"""GPU parity of K11 (input pre-stage) through the C ABI: bit-exact against the goldens of the reference's
`_depth_to_pcl` / `_sample_points` / `PC_sample` and against the numpy oracle on larger seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import prestage_oracle as po

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_depth_to_cloud_and_sampling_match_the_loader_goldens(cuda, golden):
    from hspose_b200 import pc_sample
    g = golden("prestage")
    B, n = g["depth"].shape[0], int(g["n_pts"])
    cloud, count = pc_sample.depth_to_pcl(_t(g["depth"], cuda), _t(g["camK64"], cuda), _t(g["xymap"], cuda),
                                          _t(g["mask"], cuda))
    cnt = count.cpu().numpy()
    ids = np.zeros((B, n), dtype=np.int32)
    for b in range(B):
        ref = g[f"pcl_{b}"]
        assert cnt[b] == ref.shape[0]
        assert np.array_equal(cloud[b, :cnt[b]].cpu().numpy(), ref)        # bit-exact, raster order
        ids[b] = g[f"ids_{b}"] if f"ids_{b}" in g else np.arange(n) % cnt[b]
    out = pc_sample.sample_points(cloud, count, n, ids=_t(ids, cuda)).cpu().numpy()
    for b in range(B):
        assert np.array_equal(out[b], g[f"sampled_{b}"])
    # device rule: the tile branch is the reference's own; the random-subset branch returns n DISTINCT cloud rows
    status = torch.zeros(1, dtype=torch.int32, device=cuda)
    import hspose_b200.ops as ops
    dev_out = ops.sample_points(cloud, count, n, seed=7, status=status).cpu().numpy()
    assert int(status) == 0
    for b in range(B):
        if cnt[b] <= n:
            assert np.array_equal(dev_out[b], g[f"sampled_{b}"])
        else:
            ref = g[f"pcl_{b}"]
            rows = {r.tobytes(): i for i, r in enumerate(ref)}
            picked = [rows[r.tobytes()] for r in dev_out[b]]
            assert len(set(picked)) == n and picked == sorted(picked)


def test_PC_sample_matches_the_reference_with_the_same_numpy_seed(cuda, golden):
    from hspose_b200 import pc_sample
    g = golden("prestage")
    np.random.seed(321)
    PC = pc_sample.PC_sample(_t(g["mask"], cuda)[:, None], _t(g["depth"], cuda)[:, None],
                             _t(g["camK64"].astype(np.float32), cuda), _t(g["xymap"], cuda))
    assert np.array_equal(PC.cpu().numpy(), g["PC_sample"])


def test_HSPose_forward_from_depth(cuda, golden):
    """The `depth=` path of HSPose.forward (reference HSPose.py:40-48) feeds the same cloud as PC=..."""
    from hspose_b200.HSPose import HSPose
    from hspose_b200.flags import FLAGS
    g = golden("prestage")
    B = g["depth"].shape[0]
    torch.manual_seed(0)
    model = HSPose("PoseNet_only").to(cuda).eval()
    obj_id = torch.zeros(B, dtype=torch.int64, device=cuda)
    prev = FLAGS.train
    FLAGS.train = 0
    try:
        with torch.no_grad():
            torch.manual_seed(1)
            a = model(PC=_t(g["PC_sample"], cuda), obj_id=obj_id)
            np.random.seed(321)
            torch.manual_seed(1)
            b = model(depth=_t(g["depth"], cuda)[:, None], def_mask=_t(g["mask"], cuda)[:, None],
                      camK=_t(g["camK64"].astype(np.float32), cuda), gt_2D=_t(g["xymap"], cuda), obj_id=obj_id)
    finally:
        FLAGS.train = prev
    assert torch.equal(a["PC"], b["PC"])
    assert torch.allclose(a["Pred_T"], b["Pred_T"], atol=1e-6)


@pytest.mark.parametrize("H,W,B", [(256, 256, 16), (37, 53, 3), (32, 32, 2)])
def test_depth_to_cloud_vs_oracle_ragged_and_full_size(cuda, H, W, B):
    """Full ROI size of the reference loader (256 x 256) and ragged sizes; empty and full masks included."""
    import hspose_b200.ops as ops
    rng = np.random.RandomState(H * 7 + W)
    depth = rng.randint(0, 3000, size=(B, H, W)).astype(np.float32)
    depth[rng.rand(B, H, W) < 0.3] = 0
    mask = (rng.rand(B, H, W) < 0.5).astype(np.float32)
    mask[0] = 1.0
    mask[-1] = 0.0                      # an object without any valid pixel
    xymap = rng.rand(B, 2, H, W).astype(np.float32) * 640
    camK = np.tile(np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]]), (B, 1, 1))
    n = 1028
    for K in (camK, camK.astype(np.float32)):          # float64 (loader) and float32 (PC_sample) arithmetic
        cloud, count = ops.depth_to_cloud(_t(depth, cuda), _t(mask, cuda), _t(xymap, cuda), _t(K, cuda))
        cnt = count.cpu().numpy()
        for b in range(B):
            if K.dtype == np.float64:
                ref = po.depth_to_pcl(depth[b], K[b], xymap[b], mask[b])
            else:
                n_valid = int(((mask[b] * (depth[b] > 0)) > 0).sum())
                ref = po.pc_sample(mask[b], depth[b], K[b], xymap[b], np.arange(n_valid))
            assert cnt[b] == ref.shape[0]
            assert np.array_equal(cloud[b, :cnt[b]].cpu().numpy(), ref)
        status = torch.zeros(1, dtype=torch.int32, device=cuda)
        out = ops.sample_points(cloud, count, n, seed=3, status=status)
        assert int(status) == 1 and float(out[-1].abs().max()) == 0.0      # the empty object is flagged, zeros
        # round trip: every sampled row is a row of the compacted cloud; long clouds give distinct rows
        for b in range(B - 1):
            ref = cloud[b, :cnt[b]].cpu().numpy()
            rows = {r.tobytes() for r in ref}
            got = out[b].cpu().numpy()
            assert all(r.tobytes() in rows for r in got)
    # different seeds give different subsets, the same seed the same subset
    a = ops.sample_points(cloud, count, n, seed=11)
    b2 = ops.sample_points(cloud, count, n, seed=11)
    c = ops.sample_points(cloud, count, n, seed=12)
    assert torch.equal(a, b2)
    if cnt[0] > n:
        assert not torch.equal(a[0], c[0])
