"""ctypes/numpy front-end of oracle/hsp_oracle.c — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this.  The product package (hs-pose_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhsp_oracle.so")

DIST_NEIGHBOR = 0
DIST_NEAREST = 1


def build(force=False):
    """Compile hsp_oracle.c with gcc (no OpenMP in this image; scalar, 1 thread)."""
    src = os.path.join(_HERE, "hsp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-mfma",
                               "-ffp-contract=off", "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def knn(query, cand, k, drop_first=1, formula=DIST_NEIGHBOR):
    """(B,M,D),(B,N,D) -> (B,M,k) int64; gcn3d.py:15-36."""
    query, cand = _f32(query), _f32(cand)
    B, M, D = query.shape
    N = cand.shape[1]
    out = np.empty((B, M, k), dtype=np.int64)
    lib().orc_knn(_p(query), _p(cand), B, M, N, D, k, drop_first, formula, _p(out))
    return out


def neighbor_index(v, k):
    return knn(v, v, k, 1, DIST_NEIGHBOR)


def nearest_index(target, source):
    return knn(target, source, 1, 0, DIST_NEAREST)


def direction_norm(xyz, idx):
    xyz, idx = _f32(xyz), _i32(idx)
    B, N, k = idx.shape
    out = np.empty((B, N, k, 3), dtype=np.float32)
    lib().orc_direction_norm(_p(xyz), _p(idx), B, N, k, _p(out))
    return out


def surface_conv_fwd(xyz, idx, dirn, S, C):
    xyz, idx, dirn = _f32(xyz), _i32(idx), _f32(dirn)
    B, N, k = idx.shape
    out = np.empty((B, N, C), dtype=np.float32)
    lib().orc_surface_conv_fwd(_p(xyz), _p(idx), _p(dirn), B, N, k, S, C, _p(out))
    return out


def graph_conv_fwd(xyz, idx, dirn, P, S, C, want_argmax=False):
    xyz, idx, dirn, P = _f32(xyz), _i32(idx), _f32(dirn), _f32(P)
    B, N, k = idx.shape
    out = np.empty((B, N, C), dtype=np.float32)
    am = np.empty((B, N, S * C), dtype=np.uint8) if want_argmax else None
    lib().orc_graph_conv_fwd(_p(xyz), _p(idx), _p(dirn), _p(P), B, N, k, S, C, _p(out),
                             _p(am) if want_argmax else None)
    return (out, am) if want_argmax else out


def gather_max_fwd(feat, idx, rows=None, kuse=None):
    feat, idx = _f32(feat), _i32(idx)
    B, N, C = feat.shape
    kstride = idx.shape[2]
    kuse = kstride if kuse is None else kuse
    if rows is not None:
        rows = _i32(rows)
        R = rows.shape[0]
    else:
        R = N
    out = np.empty((B, R, C), dtype=np.float32)
    lib().orc_gather_max_fwd(_p(feat), _p(idx), _p(rows) if rows is not None else None,
                             B, N, C, R, kuse, kstride, _p(out))
    return out


def orl_global_fwd(feat, idx):
    feat, idx = _f32(feat), _i32(idx)
    B, N, C = feat.shape
    k = idx.shape[2]
    G = np.empty((B, C), dtype=np.float32)
    lib().orc_orl_global_fwd(_p(feat), _p(idx), B, N, C, k, _p(G))
    return G


def upsample_rows_fwd(feat, nn, out, col0):
    feat, nn = _f32(feat), _i32(nn)
    B, Nsrc, C = feat.shape
    M = nn.shape[1]
    assert out.dtype == np.float32 and out.flags.c_contiguous
    lib().orc_upsample_rows_fwd(_p(feat), _p(nn), B, Nsrc, M, C, _p(out), out.shape[2], col0)
    return out


def chamfer_nn(a, b):
    a, b = _f32(a), _f32(b)
    B, N, _ = a.shape
    M = b.shape[1]
    dist = np.empty((B, N), dtype=np.float32)
    idx = np.empty((B, N), dtype=np.int32)
    lib().orc_chamfer_nn(_p(a), _p(b), B, N, M, _p(dist), _p(idx))
    return dist, idx
