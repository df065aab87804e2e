"""Import the UNMODIFIED reference (staged under oracle/_ref by oracle/make_ref.sh, or the live checkout)
— TEST / BASELINE INFRASTRUCTURE: only tests/, tests/golden/make_golden.py and bench.py's reference arms
use this; the product never does.

The reference imports a few heavy third-party modules it does not use on this path (matplotlib, mmcv,
detectron2, termcolor, open3d — SURVEY.md Appendix D); they are stubbed in sys.modules before the import.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    staged = os.path.join(HERE, "_ref")
    if os.path.isdir(os.path.join(staged, "network", "fs_net_repo")):
        return staged, "oracle/_ref"
    live = os.environ.get("HSPOSE_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(live, "network", "fs_net_repo")):
        return live, live
    return None, None


def import_reference(train=1):
    """-> (FLAGS, module network.HSPose, module tools.torch_utils.solver.ranger2020) or raises ImportError."""
    root, where = reference_root()
    if root is None:
        raise ImportError("reference tree not available (run oracle/make_ref.sh where /root/reference exists)")
    if root not in sys.path:
        sys.path.insert(0, root)
    argv, sys.argv = sys.argv, ["x"]

    def stub(name, **attrs):
        if name not in sys.modules:
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
    stub("open3d")
    try:
        import absl.flags as flags
        import config.config  # noqa: F401  (defines the flags)
        if not flags.FLAGS.is_parsed():
            flags.FLAGS(sys.argv)
        flags.FLAGS.train = train
        import network.HSPose as ref_hspose
        from tools.torch_utils.solver import ranger2020
    finally:
        sys.argv = argv
    return flags.FLAGS, ref_hspose, ranger2020, where
