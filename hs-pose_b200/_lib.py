"""ctypes binding of libhspose_b200.so (the C ABI in include/hspose_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhspose_b200.so")

c_int, c_void_p, c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
P = c_void_p

# name -> (restype, argtypes); must list every symbol the header declares.
SIGNATURES = {
    "hsp_version": (c_int, []),
    "hsp_strerror": (ctypes.c_char_p, [c_int]),
    "hsp_device_check": (c_int, []),
    "hsp_knn3": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "hsp_knn_feat_workspace_bytes": (c_size_t, [c_int, c_int]),
    "hsp_knn_feat": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "hsp_neighbor_direction_norm": (c_int, [P, P, c_int, c_int, c_int, P, P, P]),
    "hsp_surface_conv_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "hsp_surface_conv_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "hsp_surface_conv_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P,
                                     c_size_t, P]),
    "hsp_graph_conv_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "hsp_graph_conv_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "hsp_graph_conv_bwd": (c_int, [P, P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P, P,
                                   P, P, c_size_t, P]),
    "hsp_graph_conv_bwd_obj_supported": (c_int, [c_int] * 3),
    "hsp_graph_conv_bwd_obj_workspace_bytes": (c_size_t, [c_int] * 5),
    "hsp_graph_conv_bwd_obj": (c_int, [P, P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P, c_int, P,
                                       P, P, c_size_t, P]),
    "hsp_gather_max_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "hsp_gather_max_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "hsp_orl_global_workspace_bytes": (c_size_t, [c_int] * 3),
    "hsp_orl_global_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "hsp_orl_global_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P]),
    "hsp_upsample_rows_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, P]),
    "hsp_upsample_rows_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "hsp_bn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "hsp_bn_relu_fwd": (c_int, [P, c_int, c_int, c_int, c_int, P, P, ctypes.c_float, ctypes.c_float,
                                c_int, P, P, P, P, P, P, c_int, P, c_size_t, P]),
    "hsp_bn_relu_bwd": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, P, P, P, P, c_int, P, P, P,
                                c_int, P, P, c_size_t, P]),
    "hsp_bn_apply_fwd": (c_int, [P, c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, ctypes.c_float, ctypes.c_float,
                                 c_int, P, P, P, P, P, P, c_int, P]),
    "hsp_gemm_bf16_splits": (c_int, [c_int] * 4),
    "hsp_gemm_debug": (c_int, [c_int]),
    "hsp_gemm_bf16": (c_int, [P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, c_int,
                              c_int, c_int, P, c_int, c_int, P]),
    "hsp_gemm_bf16_acc": (c_int, [P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, c_int, P,
                                  c_int, c_int, c_int, P, c_int, c_int, P]),
    "hsp_losses_num_terms": (c_int, []),
    "hsp_losses_num_sums": (c_int, []),
    "hsp_losses_fwd": (c_int, [P, P, P, P, P, P, c_int, c_int, P, P, P]),
    "hsp_losses_bwd": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, P, P, P, P, P]),
    "hsp_optim_workspace_bytes": (c_size_t, []),
    "hsp_optim_step": (c_int, [c_int, P, P, P, P, P, ctypes.c_long, P, P, c_int, P, P, ctypes.c_float, ctypes.c_float,
                               ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_int, c_int, P, P,
                               c_size_t, P]),
    "hsp_augment": (c_int, [P] * 15 + [ctypes.c_float] * 5 + [c_int] * 3 + [P] * 5),
    "hsp_normalize_cols_fwd": (c_int, [P, c_int, ctypes.c_float, P, P, P]),
    "hsp_normalize_cols_bwd": (c_int, [P, P, P, c_int, ctypes.c_float, P, P]),
    "hsp_depth_to_cloud": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "hsp_sample_points": (c_int, [P, P, P, ctypes.c_ulonglong, c_int, c_int, c_int, P, P, P]),
    "hsp_split_bf16": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "hsp_residual_sum_fwd": (c_int, [P, P, c_int, P, P, c_int, P, P, c_int, c_int, c_int, P, P]),
    "hsp_residual_sum_bwd": (c_int, [P, P, c_int, c_int, c_int, P, P, P, P]),
    "hsp_colmax_fwd": (c_int, [P, c_int, c_int, c_int, c_int, P, P, P]),
    "hsp_chamfer_fwd": (c_int, [P, P, c_int, c_int, c_int, P, P, P, P, P]),
    "hsp_chamfer_bwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P, P, P]),
}

_lib = None


class HSPoseLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Fails loudly — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HSPoseLibraryError(
            f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
            "hs-pose_b200 has no CPU / PyTorch fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().hsp_strerror(code).decode()
        raise HSPoseLibraryError(f"{what} failed: {msg} ({code})")
