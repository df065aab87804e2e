"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/hspose_b200.h declares, the ctypes table matches the header prototype by
prototype, and the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "hspose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|size_t|const char\*)\s+(hsp_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(2)] = (m.group(1), n)
    return protos


def test_library_exports_every_header_symbol():
    from hspose_b200 import _lib
    protos = _header_prototypes()
    assert len(protos) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH)          # loads on a GPU-less host
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header():
    from hspose_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES), set(protos) ^ set(_lib.SIGNATURES)
    for name, (ret, nargs) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert len(args) == nargs, (name, len(args), nargs)
        assert (res is ctypes.c_size_t) == (ret == "size_t"), name


def test_version_and_strerror_without_gpu():
    from hspose_b200 import _lib
    lib = _lib.load()
    assert lib.hsp_version() >= 100
    assert b"argument" in lib.hsp_strerror(-1).lower() or lib.hsp_strerror(-1)
    # workspace queries are pure host arithmetic
    assert lib.hsp_knn_feat_workspace_bytes(2, 1028) >= 2 * 1028 * 4   # row norms; more for the tensor-core path
    assert lib.hsp_bn_workspace_bytes(4112, 128) > 0
    assert lib.hsp_graph_conv_bwd_workspace_bytes(2, 1028, 20, 7, 128) > 0


def test_no_cpu_fallback():
    import hspose_b200.ops as ops
    from hspose_b200 import _lib, gcn3d
    v = torch.randn(1, 64, 3)
    with pytest.raises(_lib.HSPoseLibraryError):
        ops.knn3(v, v, 8)
    with pytest.raises(_lib.HSPoseLibraryError):
        gcn3d.get_neighbor_index(v, 8)
    with pytest.raises(_lib.HSPoseLibraryError):
        gcn3d.HSlayer_surface(16, 7)(v, 8)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hs-pose_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_hspose_boundary_error_behaviour_and_build_params():
    """Reference conventions at the module boundary (network/HSPose.py:23-50,258-275,
    engine/organize_loss.py:13): unknown stages raise NotImplementedError, a missing point cloud
    without a depth ROI to sample it from is an error (and the sampling has no CPU path), build_params returns the single parameter group engine/train.py:47 expects."""
    import hspose_b200.flags as hf
    from hspose_b200.HSPose import HSPose, control_loss
    with pytest.raises(NotImplementedError):
        control_loss("no_such_stage")
    with pytest.raises(NotImplementedError):
        HSPose("no_such_stage")
    net = HSPose("PoseNet_only")
    with pytest.raises(ValueError):
        net(PC=None, obj_id=torch.zeros(1))
    from hspose_b200._lib import HSPoseLibraryError
    with pytest.raises(HSPoseLibraryError):
        net(depth=torch.ones(1, 1, 8, 8), def_mask=torch.ones(1, 1, 8, 8), camK=torch.eye(3)[None],
            gt_2D=torch.zeros(1, 2, 8, 8), obj_id=torch.zeros(1))
    groups = net.build_params(training_stage_freeze=[])
    assert len(groups) == 1 and groups[0]["lr"] == float(hf.get_flags().lr) * hf.get_flags().lr_pose
    n = sum(p.numel() for p in groups[0]["params"])
    assert n == 9709871            # train build
    frozen = HSPose("PoseNet_only").build_params(training_stage_freeze=["pose"])
    assert sum(1 for _ in frozen[0]["params"]) == 0


def test_flags_proxy_follows_the_reference_namespace_when_present():
    """FLAGS reads resolve at access time (the reference mutates FLAGS.train between construction and
    forward, evaluation/evaluate.py:39); stand-alone the reference defaults are used."""
    import hspose_b200.flags as hf
    F = hf.get_flags()
    assert F.gcn_n_num == 20 and F.gcn_sup_num == 7 and F.random_points == 1028 and F.feat_c_R == 1286
    old = F.train
    try:
        hf.FLAGS.train = 0
        assert hf.get_flags().train == 0 and hf.FLAGS.train == 0
    finally:
        hf.FLAGS.train = old


def test_argument_validation_returns_einval_without_touching_a_gpu():
    """Every entry point validates its arguments before any CUDA call: NULL pointers / bad sizes come back as
    HSP_EINVAL on a box without a device (error behaviour of the boundary, checked on the round-2 entry points)."""
    import ctypes
    from hspose_b200 import _lib
    lib = _lib.load()
    einval = lib.hsp_depth_to_cloud(None, None, None, None, 0, 1, 8, 8, None, None, None)
    assert einval != 0 and lib.hsp_strerror(einval)
    assert lib.hsp_sample_points(None, None, None, ctypes.c_ulonglong(0), 1, 64, 16, None, None, None) == einval
    assert lib.hsp_normalize_cols_fwd(None, 0, ctypes.c_float(1e-12), None, None, None) == einval
    assert lib.hsp_normalize_cols_bwd(None, None, None, 0, ctypes.c_float(1e-12), None, None) == einval
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    # a residual needs an fp32, unsplit output and N % 4 == 0
    assert lib.hsp_gemm_bf16_acc(p, 8, 0, p, 8, 0, 4, 6, 8, None, None, 0, 0, p, 8, p, 8, 1, 1, None, 0, 0, None) == einval
    assert lib.hsp_gemm_bf16_acc(p, 8, 0, p, 8, 0, 4, 8, 8, None, None, 0, 0, p, 8, p, 8, 0, 1, None, 0, 0, None) == einval
    assert lib.hsp_gather_max_bwd(None, None, None, None, 1, 8, 4, 2, 4, 4, None, None) == einval
    assert lib.hsp_upsample_rows_bwd(None, None, 1, 4, 8, 8, 8, 0, 0, None, None) == einval
    assert lib.hsp_chamfer_bwd(None, None, None, None, None, None, 1, 8, 8, None, None, None) == einval
