"""CPU: the pre-stage oracle (oracle/prestage_oracle.py) against goldens of the reference's own
`_depth_to_pcl` / `_sample_points` / `PC_sample` (tests/golden/prestage.npz)."""
import numpy as np

from oracle import prestage_oracle as po


def test_depth_to_pcl_bit_exact(golden):
    g = golden("prestage")
    for b in range(g["depth"].shape[0]):
        mine = po.depth_to_pcl(g["depth"][b], g["camK64"][b], g["xymap"][b], g["mask"][b])
        assert mine.dtype == np.float32 and np.array_equal(mine, g[f"pcl_{b}"])


def test_sample_points_bit_exact(golden):
    g = golden("prestage")
    n = int(g["n_pts"])
    seen = set()
    for b in range(g["depth"].shape[0]):
        pcl = g[f"pcl_{b}"]
        seen.add(np.sign(pcl.shape[0] - n))
        mine = po.sample_points(pcl, n, g.get(f"ids_{b}"))
        assert np.array_equal(mine, g[f"sampled_{b}"])
    assert {-1, 1} <= seen        # both the tile branch and the random-subset branch are covered


def test_pc_sample_bit_exact(golden):
    g = golden("prestage")
    for b in range(g["depth"].shape[0]):
        mine = po.pc_sample(g["mask"][b], g["depth"][b], g["camK64"][b].astype(np.float32), g["xymap"][b],
                            g["PC_choose"][b])
        assert np.array_equal(mine, g["PC_sample"][b])
