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
