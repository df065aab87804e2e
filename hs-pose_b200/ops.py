"""Torch-facing wrappers over the C ABI (include/hspose_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every op below
is one or two launches of a hand-written sm_100a kernel in libhspose_b200.so.
No op has a PyTorch/CPU fallback — a missing library or a CPU tensor raises.
"""
import contextlib
import ctypes
import os

import torch

from . import _lib

DIST_NEIGHBOR = 0
DIST_NEAREST = 1
F32, BF16 = 0, 1   # HSP_DTYPE_*

_launches = 0  # number of CUDA kernels launched through the C ABI (bench.py reads this)

# kernels launched per entry point (memsets not counted)
_KERNELS_PER_CALL = {"hsp_optim_step": 3, "hsp_losses_fwd": 2, "hsp_losses_bwd": 2, "hsp_bn_apply_fwd": 2, "hsp_knn_feat": 2, "hsp_surface_conv_bwd": 2, "hsp_graph_conv_bwd": 2,
                     "hsp_orl_global_fwd": 2, "hsp_chamfer_fwd": 2, "hsp_chamfer_bwd": 2,
                     "hsp_bn_relu_fwd": 3, "hsp_bn_relu_bwd": 3}

_timing = None  # when a list: (name, int-args, start_event, end_event) per call


def launch_count():
    return _launches


def enable_timing(on=True):
    """Record a CUDA-event pair around every C-ABI call on the launching stream."""
    global _timing
    _timing = [] if on else None


def timing_records():
    """[(entry point, integer args, milliseconds)] — call after torch.cuda.synchronize()."""
    return [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in (_timing or [])]


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.HSPoseLibraryError(
            f"{name}: expected a CUDA tensor (hs-pose_b200 has no CPU path), got "
            f"{type(t).__name__} on {getattr(t, 'device', '?')}")
    if t.dtype != dtype:
        if dtype == torch.float32 and t.dtype in (torch.bfloat16, torch.float16):
            t = t.float()   # autocast hands us half tensors; the kernels compute in fp32
        else:
            raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


_NVTX = os.environ.get("HSP_NVTX", "1") != "0"   # an NVTX range per C-ABI call (tracing hook)


def _call(name, *args):
    global _launches
    lib = _lib.load()
    _launches += _KERNELS_PER_CALL.get(name, 1)
    if _NVTX:
        torch.cuda.nvtx.range_push(name)
        try:
            return _call_inner(lib, name, args)
        finally:
            torch.cuda.nvtx.range_pop()
    return _call_inner(lib, name, args)


def _call_inner(lib, name, args):
    if _timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(getattr(lib, name)(*args), name)
        e1.record()
        _timing.append((name, tuple(a for a in args if isinstance(a, int)), e0, e1))
        return
    _lib.check(getattr(lib, name)(*args), name)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------- KNN
def knn3(query, cand, k, drop_first=1, formula=DIST_NEIGHBOR, want64=False, want32=True):
    """Fused 3-D distance + top-k (K1).  Returns (idx64 | None, idx32 | None)."""
    query = _need(query, torch.float32, "query")
    cand = _need(cand, torch.float32, "cand")
    B, M, D = query.shape
    if D != 3 or cand.shape[0] != B or cand.shape[2] != 3:
        raise ValueError("knn3 expects (B,M,3) and (B,N,3)")
    N = cand.shape[1]
    with torch.cuda.device(query.device):
        i64 = torch.empty(B, M, k, dtype=torch.int64, device=query.device) if want64 else None
        i32 = torch.empty(B, M, k, dtype=torch.int32, device=query.device) if want32 else None
        _call("hsp_knn3", _p(query), _p(cand), B, M, N, k, drop_first, formula, _p(i64), _p(i32),
              _stream())
    return i64, i32


def knn_feat(feat, k, drop_first=1, want64=False, want32=True):
    """Fused D-dim (feature-space) distance + top-k (K2)."""
    feat = _need(feat, torch.float32, "feat")
    B, N, D = feat.shape
    with torch.cuda.device(feat.device):
        lib = _lib.load()
        ws = _workspace(lib.hsp_knn_feat_workspace_bytes(B, N), feat.device)
        i64 = torch.empty(B, N, k, dtype=torch.int64, device=feat.device) if want64 else None
        i32 = torch.empty(B, N, k, dtype=torch.int32, device=feat.device) if want32 else None
        _call("hsp_knn_feat", _p(feat), B, N, D, k, drop_first, _p(i64), _p(i32), _p(ws),
              ws.numel(), _stream())
        if N >= 128 and D in (128, 256) and k + drop_first <= 64:
            global _launches
            _launches += 2   # tensor-core path: norm, split, filter, refine (4 kernels instead of 2)
    return i64, i32


def direction_norm(xyz, idx32, return_unnormed=False):
    xyz = _need(xyz, torch.float32, "xyz")
    idx32 = _need(idx32, torch.int32, "idx")
    B, N, k = idx32.shape
    with torch.cuda.device(xyz.device):
        out = torch.empty(B, N, k, 3, dtype=torch.float32, device=xyz.device)
        raw = torch.empty_like(out) if return_unnormed else None
        _call("hsp_neighbor_direction_norm", _p(xyz), _p(idx32), B, N, k, _p(out), _p(raw),
              _stream())
    return (out, raw) if return_unnormed else out


# ------------------------------------------------------------ graph convs
def _no_xyz_grad(ctx):
    """The reference graph is differentiable w.r.t. the vertices through F.normalize(neighbours - vertices); the
    fused kernels do not produce that gradient (the network's input cloud is detached, HSPose.py:53).  A caller
    whose coordinates require grad gets an error instead of a silently missing gradient."""
    if ctx.needs_input_grad[0]:
        raise NotImplementedError("hs-pose_b200: the gradient w.r.t. the point coordinates (xyz) is not implemented; "
                                  "detach the cloud (the reference detaches it, network/HSPose.py:53)")


class _NormalizeCols(torch.autograd.Function):
    """F.normalize(d, dim=0) for the (3, S*C) support directions, one launch each way."""

    @staticmethod
    def forward(ctx, d):
        d = _need(d, torch.float32, "directions")
        n = d.shape[1]
        with torch.cuda.device(d.device):
            out, nrm = torch.empty_like(d), torch.empty(n, dtype=torch.float32, device=d.device)
            _call("hsp_normalize_cols_fwd", _p(d), n, ctypes.c_float(1e-12), _p(out), _p(nrm), _stream())
        ctx.save_for_backward(out, nrm)
        return out

    @staticmethod
    def backward(ctx, g):
        out, nrm = ctx.saved_tensors
        g = _need(g, torch.float32, "g")
        with torch.cuda.device(g.device):
            gd = torch.empty_like(out)
            _call("hsp_normalize_cols_bwd", _p(g), _p(out), _p(nrm), out.shape[1], ctypes.c_float(1e-12), _p(gd),
                  _stream())
        return gd


def normalize_dirs(d):
    """Unit support directions (reference gcn3d.py:95, :162: F.normalize(self.directions, dim=0))."""
    if d.dim() != 2 or d.shape[0] != 3:
        raise ValueError("normalize_dirs expects the (3, S*C) direction parameter")
    return _NormalizeCols.apply(d)


class _SplitHalves(torch.autograd.Function):
    """(W[:, :h], W[:, h:]) of the ORL 1x1 convolution weight (C, 2C); the gradient is ONE concatenation instead of
    two zero-filled (C, 2C) buffers, two slice copies and an add."""

    @staticmethod
    def forward(ctx, W, h):
        ctx.h, ctx.shape = h, W.shape
        return W[:, :h], W[:, h:]

    @staticmethod
    def backward(ctx, ga, gb):
        if ga is None:
            ga = gb.new_zeros(ctx.shape[0], ctx.h)
        if gb is None:
            gb = ga.new_zeros(ctx.shape[0], ctx.shape[1] - ctx.h)
        return torch.cat([ga, gb], dim=1), None


def split_halves(W, h):
    return _SplitHalves.apply(W, h)


class _SurfaceConv(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, xyz, idx32, dirn, S, C):
        _no_xyz_grad(ctx)
        xyz = _need(xyz, torch.float32, "xyz")
        idx32 = _need(idx32, torch.int32, "idx")
        dirn = _need(dirn, torch.float32, "dirn")
        B, N, k = idx32.shape
        need_grad = ctx.needs_input_grad[2]
        with torch.cuda.device(xyz.device):
            out = torch.empty(B, N, C, dtype=torch.float32, device=xyz.device)
            am = torch.empty(B, N, S * C, dtype=torch.uint8, device=xyz.device) if need_grad else None
            _call("hsp_surface_conv_fwd", _p(xyz), _p(idx32), _p(dirn), B, N, k, S, C, _p(out),
                  _p(am), _stream())
        if need_grad:
            ctx.save_for_backward(xyz, idx32, am)
        ctx.dims = (B, N, k, S, C)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        xyz, idx32, am = ctx.saved_tensors
        B, N, k, S, C = ctx.dims
        gout = _need(gout.float(), torch.float32, "gout")
        with torch.cuda.device(xyz.device):
            lib = _lib.load()
            ws = _workspace(lib.hsp_surface_conv_bwd_workspace_bytes(B, N, k, S, C), xyz.device)
            gdirn = torch.empty(3, S * C, dtype=torch.float32, device=xyz.device)
            _call("hsp_surface_conv_bwd", _p(xyz), _p(idx32), _p(am), _p(gout), B, N, k, S, C,
                  _p(gdirn), _p(ws), ws.numel(), _stream())
        return None, None, gdirn, None, None


def surface_conv(xyz, idx32, dirn, S, C):
    """K3: out (B,N,C) = mean_s max_n relu(rhat . dirn)."""
    return _SurfaceConv.apply(xyz, idx32, dirn, S, C)


def _graph_conv_fwd_raw(xyz, idx32, dirn, P, S, C, want_argmax):
    B, N, k = idx32.shape
    if P.shape != (B, N, (S + 1) * C):
        raise ValueError(f"P must be (B,N,(S+1)*C), got {tuple(P.shape)}")
    dt = BF16 if P.dtype == torch.bfloat16 else F32
    with torch.cuda.device(xyz.device):
        out = torch.empty(B, N, C, dtype=torch.float32, device=xyz.device)
        am = torch.empty(B, N, S * C, dtype=torch.uint8, device=xyz.device) if want_argmax else None
        _call("hsp_graph_conv_fwd", _p(xyz), _p(idx32), _p(dirn), _p(P), dt, B, N, k, S, C,
              _p(out), _p(am), _stream())
    return out, am


def _graph_conv_bwd_raw(xyz, idx32, dirn, P, am, gout, S, C, want_gbias=False, gp_dtype=torch.float32, variant="auto"):
    """K4b.  variant "obj" (object-resident slabs, gP written once in gp_dtype) when the shape allows it,
    else "atomic" (fp32 gP via global float atomics; a bf16 gP is then one cast pass)."""
    B, N, k = idx32.shape
    dt = BF16 if P.dtype == torch.bfloat16 else F32
    with torch.cuda.device(xyz.device):
        lib = _lib.load()
        gdirn = torch.empty_like(dirn)
        gbias = torch.empty((S + 1) * C, dtype=torch.float32, device=xyz.device) if want_gbias else None
        use_obj = S == 7 and lib.hsp_graph_conv_bwd_obj_supported(N, k, C) == 1 if variant == "auto" else variant == "obj"
        if use_obj:
            ws = _workspace(lib.hsp_graph_conv_bwd_obj_workspace_bytes(B, N, k, S, C), xyz.device)
            gP = torch.empty(B, N, (S + 1) * C, dtype=gp_dtype, device=xyz.device)
            _call("hsp_graph_conv_bwd_obj", _p(xyz), _p(idx32), _p(dirn), _p(P), dt, _p(am), _p(gout), B, N, k, S, C,
                  _p(gP), BF16 if gp_dtype == torch.bfloat16 else F32, _p(gdirn), _p(gbias), _p(ws), ws.numel(),
                  _stream())
        else:
            ws = _workspace(lib.hsp_graph_conv_bwd_workspace_bytes(B, N, k, S, C), xyz.device)
            gP = torch.empty(B, N, (S + 1) * C, dtype=torch.float32, device=xyz.device)
            _call("hsp_graph_conv_bwd", _p(xyz), _p(idx32), _p(dirn), _p(P), dt, _p(am), _p(gout),
                  B, N, k, S, C, _p(gP), _p(gdirn), _p(gbias), _p(ws), ws.numel(), _stream())
            if gp_dtype != torch.float32:
                gP = gP.to(gp_dtype)
    return (gP, gdirn, gbias) if want_gbias else (gP, gdirn)


class _GraphConv(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, xyz, idx32, dirn, P, S, C):
        _no_xyz_grad(ctx)
        xyz = _need(xyz, torch.float32, "xyz")
        idx32 = _need(idx32, torch.int32, "idx")
        dirn = _need(dirn, torch.float32, "dirn")
        P = _need(P, torch.float32, "P")
        need_grad = any(ctx.needs_input_grad)
        out, am = _graph_conv_fwd_raw(xyz, idx32, dirn, P, S, C, need_grad)
        if need_grad:
            ctx.save_for_backward(xyz, idx32, dirn, P, am)
        ctx.dims = (S, C)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        xyz, idx32, dirn, P, am = ctx.saved_tensors
        S, C = ctx.dims
        gout = _need(gout.float(), torch.float32, "gout")
        gP, gdirn = _graph_conv_bwd_raw(xyz, idx32, dirn, P, am, gout, S, C)
        return None, None, gdirn, gP, None, None


def graph_conv(xyz, idx32, dirn, P, S, C):
    """K4: out = P[..., :C] + mean_s max_n(relu(rhat . dirn) * P[idx, C + s*C + c])."""
    return _GraphConv.apply(xyz, idx32, dirn, P, S, C)


class _HSConvMixed(torch.autograd.Function):
    """Mixed-precision HS graph convolution: P = fm @ W + b is produced in bf16 by a
    tensor-core GEMM and consumed in bf16 by the fused gather kernel (half the gather
    traffic, no fp32 copy of P); the backward kernel emits gP in fp32 (atomics) and the
    two weight/input-gradient GEMMs run on it directly in TF32 — no cast passes over the
    (B,N,(S+1)C) tensors.  One autograd node, so the fp32 gP never gets down-cast."""

    @staticmethod
    def forward(ctx, xyz, idx32, dirn, fm, W, bias, S, C):
        _no_xyz_grad(ctx)
        xyz = _need(xyz, torch.float32, "xyz")
        idx32 = _need(idx32, torch.int32, "idx")
        dirn = _need(dirn, torch.float32, "dirn")
        fm = _need(fm, torch.float32, "feature_map")
        B, N, Cin = fm.shape
        fm16 = fm.reshape(B * N, Cin).to(torch.bfloat16)
        W16 = _as_gemm_operand(W)                        # (Cin, (S+1)C): read MN-major as the B operand
        P = gemm_bf16(fm16, W16, b_mn=True, bias=bias).view(B, N, (S + 1) * C)
        need_grad = any(ctx.needs_input_grad)
        out, am = _graph_conv_fwd_raw(xyz, idx32, dirn, P, S, C, need_grad)
        if need_grad:
            ctx.save_for_backward(xyz, idx32, dirn, P, am, fm16, W16)
        ctx.dims = (S, C)
        return out

    @staticmethod
    def backward(ctx, gout):
        xyz, idx32, dirn, P, am, fm16, W16 = ctx.saved_tensors
        S, C = ctx.dims
        B, N = idx32.shape[0], idx32.shape[1]
        Cin = fm16.shape[1]
        gout = _need(gout.float(), torch.float32, "gout")
        # weight / input gradients on the K6 kernel (bf16 operands, fp32 accumulate): K4b emits gP in bf16
        gP16, gdirn, gb = _graph_conv_bwd_raw(xyz, idx32, dirn, P, am, gout, S, C, want_gbias=True,
                                              gp_dtype=torch.bfloat16)
        gP16 = gP16.view(B * N, (S + 1) * C)
        # the input gradient leaves the GEMM epilogue in fp32 (feature maps are fp32: no bf16 round trip + cast pass)
        gfm = (gemm_bf16(gP16, W16, out_dtype=torch.float32).view(B, N, Cin)
               if ctx.needs_input_grad[3] else None)
        gW = gemm_bf16(fm16, gP16, a_mn=True, b_mn=True, out_dtype=torch.float32,
                       splits=gemm_splits(Cin, (S + 1) * C, B * N))
        return None, None, gdirn, gfm, gW, gb, None, None


def hs_conv_mixed(xyz, idx32, dirn, fm, W, bias, S, C):
    return _HSConvMixed.apply(xyz, idx32, dirn, fm, W, bias, S, C)


# ---------------------------------------------------------------- gathers
class _GatherMax(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, feat, idx32, rows32, kuse):
        feat = _need(feat, torch.float32, "feat")
        idx32 = _need(idx32, torch.int32, "idx")
        B, N, C = feat.shape
        kstride = idx32.shape[2]
        R = N if rows32 is None else rows32.numel()
        if rows32 is not None:
            rows32 = _need(rows32, torch.int32, "rows")
        need_grad = ctx.needs_input_grad[0]
        with torch.cuda.device(feat.device):
            out = torch.empty(B, R, C, dtype=torch.float32, device=feat.device)
            am = torch.empty(B, R, C, dtype=torch.uint8, device=feat.device) if need_grad else None
            _call("hsp_gather_max_fwd", _p(feat), _p(idx32), _p(rows32), B, N, C, R, kuse,
                  kstride, _p(out), _p(am), _stream())
        if need_grad:
            ctx.save_for_backward(idx32, rows32, am)
        ctx.dims = (B, N, C, R, kuse, kstride)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        idx32, rows32, am = ctx.saved_tensors
        B, N, C, R, kuse, kstride = ctx.dims
        gout = _need(gout.float(), torch.float32, "gout")
        with torch.cuda.device(gout.device):
            gfeat = torch.empty(B, N, C, dtype=torch.float32, device=gout.device)   # written in full by the call
            _call("hsp_gather_max_bwd", _p(gout), _p(idx32), _p(rows32), _p(am), B, N, C, R,
                  kuse, kstride, _p(gfeat), _stream())
        return gfeat, None, None, None


def gather_max(feat, idx32, rows32=None, kuse=None):
    """K5b: out[b,r,c] = max_{n<kuse} feat[b, idx[b, rows[r], n], c]."""
    return _GatherMax.apply(feat, idx32, rows32, idx32.shape[2] if kuse is None else kuse)


class _OrlGlobal(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, feat, idx32):
        feat = _need(feat, torch.float32, "feat")
        idx32 = _need(idx32, torch.int32, "idx")
        B, N, C = feat.shape
        k = idx32.shape[2]
        need_grad = ctx.needs_input_grad[0]
        with torch.cuda.device(feat.device):
            lib = _lib.load()
            ws = _workspace(lib.hsp_orl_global_workspace_bytes(B, N, C), feat.device)
            G = torch.empty(B, C, dtype=torch.float32, device=feat.device)
            am = torch.empty(B, N, C, dtype=torch.uint8, device=feat.device) if need_grad else None
            _call("hsp_orl_global_fwd", _p(feat), _p(idx32), B, N, C, k, _p(G), _p(am), _p(ws),
                  ws.numel(), _stream())
        if need_grad:
            ctx.save_for_backward(idx32, am)
        ctx.dims = (B, N, C, k)
        return G

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gG):
        idx32, am = ctx.saved_tensors
        B, N, C, k = ctx.dims
        gG = _need(gG.float(), torch.float32, "gG")
        with torch.cuda.device(gG.device):
            gfeat = torch.zeros(B, N, C, dtype=torch.float32, device=gG.device)
            _call("hsp_orl_global_bwd", _p(gG), _p(idx32), _p(am), B, N, C, k, _p(gfeat),
                  _stream())
        return gfeat, None


def orl_global(feat, idx32):
    """K5a: G (B,C) = mean_i max_n feat[b, idx[b,i,n], c]."""
    return _OrlGlobal.apply(feat, idx32)


class _GatherRows(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, feat, nn32):
        feat = _need(feat, torch.float32, "feat")
        nn32 = _need(nn32, torch.int32, "nn")
        B, Nsrc, C = feat.shape
        M = nn32.shape[1]
        with torch.cuda.device(feat.device):
            out = torch.empty(B, M, C, dtype=torch.float32, device=feat.device)
            _call("hsp_upsample_rows_fwd", _p(feat), _p(nn32), B, Nsrc, M, C, _p(out), C, 0, F32,
                  _stream())
        ctx.save_for_backward(nn32)
        ctx.dims = (B, Nsrc, M, C)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        (nn32,) = ctx.saved_tensors
        B, Nsrc, M, C = ctx.dims
        gout = _need(gout.float(), torch.float32, "gout")
        with torch.cuda.device(gout.device):
            gfeat = torch.empty(B, Nsrc, C, dtype=torch.float32, device=gout.device)   # written in full by the call
            _call("hsp_upsample_rows_bwd", _p(gout), _p(nn32), B, Nsrc, M, C, C, 0, F32, _p(gfeat),
                  _stream())
        return gfeat, None


def gather_rows(feat, nn32):
    """K5c: out[b,i,:] = feat[b, nn[b,i], :]   (nn (B,M) int32)."""
    return _GatherRows.apply(feat, nn32)


class _ConcatUpsample(torch.autograd.Function):
    """feat (B,M,ld) = cat_i rows(piece_i) [+ zero padding up to ld]: piece_i is copied as
    is (nn None, Nsrc == M), gathered through a nearest-neighbour table nn_i (B,M) int32,
    or broadcast over points ((B,C) piece).  One launch per piece, straight into the
    concat buffer (FaceRecon.py:100-107 without the intermediate tensors); the buffer
    may be bf16 and padded to an aligned row length for the tensor-core MLPs."""

    @staticmethod
    def forward(ctx, M, nns, ld, out_dtype, *pieces):
        pieces = [_need(p, torch.float32, "piece") for p in pieces]
        B = pieces[0].shape[0]
        widths = [p.shape[-1] for p in pieces]
        tot = sum(widths)
        ld = tot if ld is None else ld
        dev = pieces[0].device
        dt = BF16 if out_dtype == torch.bfloat16 else F32
        nn32 = []
        with torch.cuda.device(dev):
            out = torch.empty(B, M, ld, dtype=out_dtype, device=dev)
            if ld > tot:
                out[:, :, tot:].zero_()
            col = 0
            for p, nn, w in zip(pieces, nns, widths):
                if isinstance(nn, str):  # "bcast": (B,C) -> every point
                    nsrc, t = 1, None
                elif nn is None:
                    nsrc, t = p.shape[1], None
                    if nsrc != M:
                        raise ValueError("identity piece must have M rows")
                else:
                    nsrc, t = p.shape[1], _need(nn, torch.int32, "nn")
                nn32.append(t)
                _call("hsp_upsample_rows_fwd", _p(p), _p(t), B, nsrc, M, w, _p(out), ld, col, dt,
                      _stream())
                col += w
        ctx.save_for_backward(*[t for t in nn32 if t is not None])
        ctx.meta = (B, M, ld, dt, widths, [t is not None for t in nn32],
                    [isinstance(nn, str) for nn in nns], [p.shape[1] if p.dim() == 3 else 1 for p in pieces])
        return out

    @staticmethod
    def backward(ctx, gout):
        B, M, ld, dt, widths, has_nn, bcast, nsrcs = ctx.meta
        saved = list(ctx.saved_tensors)
        gout = _need(gout, torch.bfloat16 if dt == BF16 else torch.float32, "gout")
        grads, col = [], 0
        with torch.cuda.device(gout.device):
            for i, w in enumerate(widths):
                nn = saved.pop(0) if has_nn[i] else None
                if not ctx.needs_input_grad[4 + i]:
                    grads.append(None)
                elif bcast[i]:
                    grads.append(gout[:, :, col:col + w].sum(dim=1, dtype=torch.float32))
                else:
                    g = torch.empty(B, nsrcs[i] if nn is not None else M, w, dtype=torch.float32, device=gout.device)
                    _call("hsp_upsample_rows_bwd", _p(gout), _p(nn), B, nsrcs[i], M, w, ld, col, dt,
                          _p(g), _stream())
                    grads.append(g)
                col += w
        return (None, None, None, None, *grads)


def concat_upsample(pieces, nns, M, ld=None, out_dtype=torch.float32):
    """K5c + concat.  pieces[i]: (B,M,C) with nns[i] None, (B,Nsrc,C) with nns[i] a
    (B,M) int32 nearest table, or (B,C) with nns[i] == "bcast".  Columns past the
    pieces (up to ld) are zero."""
    return _ConcatUpsample.apply(M, list(nns), ld, out_dtype, *pieces)


class _ResidualSum(torch.autograd.Function):
    """out = feature + lin + gproj[:, None, :] + ste in one pass (K5d); `ste` is either a (B,N,C)
    tensor or, for the surface layer, the pair (xyz (B,N,3), wxyz (C,3)) evaluated in place.
    Backward = one pass that emits the bf16 cast of the gradient (for bf16 lin / ste), the per-object
    column sums and the per-object partials of d wxyz."""

    @staticmethod
    def forward(ctx, feature, lin, gproj, ste, xyz, wxyz):
        feature = _need(feature, torch.float32, "feature")
        B, N, C = feature.shape

        def prep(t, name):
            if t is None:
                return None, F32
            if t.dtype not in (torch.float32, torch.bfloat16) or t.shape != feature.shape:
                raise TypeError(f"{name}: expected (B,N,C) fp32/bf16")
            return (t if t.is_contiguous() else t.contiguous()), (BF16 if t.dtype == torch.bfloat16 else F32)
        lin, ldt = prep(lin, "lin")
        ste, sdt = prep(ste, "ste")
        gproj = _need(gproj, torch.float32, "gproj") if gproj is not None else None
        if xyz is not None:
            xyz = _need(xyz, torch.float32, "xyz")
            wxyz = _need(wxyz, torch.float32, "wxyz")
            if xyz.shape != (B, N, 3) or wxyz.shape != (C, 3) or ste is not None:
                raise ValueError("residual_sum: xyz (B,N,3) + wxyz (C,3) replace ste")
        with torch.cuda.device(feature.device):
            out = torch.empty_like(feature)
            _call("hsp_residual_sum_fwd", _p(feature), _p(lin), ldt, _p(gproj), _p(ste), sdt, _p(xyz),
                  _p(wxyz), B, N, C, _p(out), _stream())
        ctx.save_for_backward(xyz)
        ctx.meta = (B, N, C, None if lin is None else ldt, None if ste is None else sdt, gproj is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        B, N, C, ldt, sdt, has_gp = ctx.meta
        (xyz,) = ctx.saved_tensors
        g = _need(g.float(), torch.float32, "g")
        want16 = (ldt == BF16 and ctx.needs_input_grad[1]) or (sdt == BF16 and ctx.needs_input_grad[3])
        want_gp = has_gp and ctx.needs_input_grad[2]
        want_w = xyz is not None and ctx.needs_input_grad[5]
        g16 = ggp = gwp = None
        if want16 or want_gp or want_w:
            with torch.cuda.device(g.device):
                g16 = torch.empty(B, N, C, dtype=torch.bfloat16, device=g.device) if want16 else None
                ggp = torch.empty(B, C, dtype=torch.float32, device=g.device) if want_gp else None
                gwp = torch.empty(B, C, 3, dtype=torch.float32, device=g.device) if want_w else None
                _call("hsp_residual_sum_bwd", _p(g), _p(xyz) if want_w else None, B, N, C, _p(g16), _p(ggp),
                      _p(gwp), _stream())

        def pick(dt, need):
            if dt is None or not need:
                return None
            return g16 if dt == BF16 else g
        return (g if ctx.needs_input_grad[0] else None, pick(ldt, ctx.needs_input_grad[1]), ggp,
                pick(sdt, ctx.needs_input_grad[3]), None, gwp.sum(dim=0) if want_w else None)


def residual_sum(feature, lin=None, gproj=None, ste=None, xyz=None, wxyz=None):
    """K5d: feature + lin + gproj[:, None, :] + ste  ((B,N,C); gproj (B,C)); ste may be given as
    xyz (B,N,3) + wxyz (C,3) (the surface layer's coordinate STE)."""
    return _ResidualSum.apply(feature, lin, gproj, ste, xyz, wxyz)


class _ColMax(torch.autograd.Function):
    """out[b,c] = max_i x[b,i,c] on a point-major (B,N,C) activation (fp32 / bf16)."""

    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
            raise _lib.HSPoseLibraryError("colmax: expected a CUDA fp32/bf16 (B,N,C) tensor")
        x = x if x.is_contiguous() else x.contiguous()
        B, N, C = x.shape
        with torch.cuda.device(x.device):
            out = torch.empty(B, C, dtype=x.dtype, device=x.device)
            arg = torch.empty(B, C, dtype=torch.int32, device=x.device)
            _call("hsp_colmax_fwd", _p(x), BF16 if x.dtype == torch.bfloat16 else F32, B, N, C, _p(out),
                  _p(arg), _stream())
        ctx.save_for_backward(arg)
        ctx.meta = (B, N, C, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        B, N, C, dt = ctx.meta
        gx = torch.zeros(B, N, C, dtype=dt, device=g.device)
        gx.scatter_(1, arg.long().unsqueeze(1), g.to(dt).unsqueeze(1))
        return gx


def colmax(x):
    """Max over the points (dim 1) of a (B,N,C) activation."""
    return _ColMax.apply(x)


# ------------------------------------------------------------ BatchNorm + ReLU
class _BnRelu(torch.autograd.Function):
    """Batch-statistics BatchNorm1d (+ReLU) on a (M,C) matrix (K6b); x may be a column
    slice (stride (ld,1)) of a wider matrix.  Updates the running statistics in place."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, relu):
        if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
            raise _lib.HSPoseLibraryError("bn_relu: expected a CUDA (M,C) matrix with unit column stride")
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"bn_relu: unsupported dtype {x.dtype}")
        gamma = gamma.float().contiguous()
        beta = beta.float().contiguous()
        y, stats = _bn_fwd_raw(x, gamma, beta, running_mean, running_var, eps, momentum, relu)
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.relu = int(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, stats = ctx.saved_tensors
        dx, dgamma, dbeta, _ = _bn_bwd_raw(x, dy, gamma, beta, stats, ctx.relu, False)
        return dx, dgamma, dbeta, None, None, None, None, None


def _bn_fwd_raw(x, gamma, beta, running_mean, running_var, eps, momentum, relu):
    """x (M,C) unit column stride -> (y (M,C) contiguous, stats (4,C) = mean, invstd, scale, shift)."""
    M, C = x.shape
    dt = BF16 if x.dtype == torch.bfloat16 else F32
    lib = _lib.load()
    with torch.cuda.device(x.device):
        y = torch.empty(M, C, dtype=x.dtype, device=x.device)
        stats = torch.empty(4, C, dtype=torch.float32, device=x.device)
        ws = _workspace(lib.hsp_bn_workspace_bytes(M, C), x.device)
        _call("hsp_bn_relu_fwd", _p(x), x.stride(0), M, C, dt, _p(gamma), _p(beta), float(eps),
              float(momentum), int(relu), _p(running_mean), _p(running_var), _p(stats[0]),
              _p(stats[1]), _p(stats[2]), _p(y), C, _p(ws), ws.numel(), _stream())
    return y, stats


def _bn_fwd_from_partials(x, partials, gamma, beta, running_mean, running_var, eps, momentum, relu):
    """Same as _bn_fwd_raw with the statistics partials (nblocks,2,C | row pitch ldp) from the GEMM epilogue."""
    M, C = x.shape
    dt = BF16 if x.dtype == torch.bfloat16 else F32
    with torch.cuda.device(x.device):
        y = torch.empty(M, C, dtype=x.dtype, device=x.device)
        stats = torch.empty(4, C, dtype=torch.float32, device=x.device)
        _call("hsp_bn_apply_fwd", _p(x), x.stride(0), M, C, dt, _p(partials), partials.shape[0],
              partials.stride(1), _p(gamma), _p(beta), float(eps), float(momentum), int(relu),
              _p(running_mean), _p(running_var), _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(y), C, _stream())
    return y, stats


def _bn_bwd_raw(x, dy, gamma, beta, stats, relu, want_colsum, dx_out=None):
    """-> (dx, dgamma, dbeta, colsum(dx) | None)."""
    M, C = x.shape
    dt = BF16 if x.dtype == torch.bfloat16 else F32
    dy = dy.to(x.dtype)
    if dy.stride(1) != 1 or (dy.stride(0) * dy.element_size()) % 16 != 0:
        dy = dy.contiguous()
    lib = _lib.load()
    with torch.cuda.device(x.device):
        dx = torch.empty(M, C, dtype=x.dtype, device=x.device) if dx_out is None else dx_out
        dgb = torch.empty(3, C, dtype=torch.float32, device=x.device)
        ws = _workspace(lib.hsp_bn_workspace_bytes(M, C), x.device)
        _call("hsp_bn_relu_bwd", _p(x), x.stride(0), _p(dy), dy.stride(0), M, C, dt, _p(gamma),
              _p(beta), _p(stats[0]), _p(stats[1]), int(relu), _p(dgb[0]), _p(dgb[1]), _p(dx), dx.stride(0),
              _p(dgb[2]) if want_colsum else None, _p(ws), ws.numel(), _stream())
    if want_colsum:
        global _launches
        _launches += 1   # the column-sum finalize
    return dx, dgb[0], dgb[1], (dgb[2] if want_colsum else None)


def bn_relu(x, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1, relu=True):
    """y = relu?(batchnorm_train(x)) over the rows of the (M,C) matrix x."""
    return _BnRelu.apply(x, gamma, beta, running_mean, running_var, eps, momentum, relu)


class _LinearBnRelu(torch.autograd.Function):
    """z = relu?(batchnorm_train(x @ W^T + b)) over the rows of x (M,K) — one Conv1d(k=1) ->
    BatchNorm1d -> ReLU block of the dense per-point MLPs (reference FaceRecon.py:38-68,
    PoseR.py:26-29, PoseTs.py:31-34) as ONE autograd node on the mixed-precision path: the K6
    tensor-core GEMM emits the BatchNorm statistics in its epilogue (no stand-alone reduction pass),
    K6b normalises; backward = K6b + the K6 dgrad / wgrad GEMMs; the Linear's bias gradient comes
    out of the BN-backward kernel (column sums of dY)."""

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, running_mean, running_var, eps, momentum, relu, bias_rows=None,
                rows_per_group=0):
        xb = _as_gemm_operand(x)
        Wb = _as_gemm_operand(W)
        y, part = gemm_bf16(xb, Wb, bias=b, stats=True, bias_rows=bias_rows, rows_per_group=rows_per_group)
        g32, b32 = gamma.float().contiguous(), beta.float().contiguous()
        z, stats = _bn_fwd_from_partials(y, part, g32, b32, running_mean, running_var, eps, momentum, relu)
        ctx.save_for_backward(xb, Wb, y, g32, b32, stats)
        ctx.relu, ctx.has_bias = int(relu), b is not None
        ctx.rpg = rows_per_group if bias_rows is not None else 0
        ctx.dx_dtype = _dx_dtype(x)
        return z

    @staticmethod
    def backward(ctx, dz):
        xb, Wb, y, g32, b32, stats = ctx.saved_tensors
        dy, dgamma, dbeta, colsum = _bn_bwd_raw(y, dz, g32, b32, stats, ctx.relu, ctx.has_bias)
        dx = gemm_bf16(dy, Wb, b_mn=True, out_dtype=ctx.dx_dtype) if ctx.needs_input_grad[0] else None
        dW = _wgrad(dy, xb)
        drows = None
        if ctx.rpg:      # d bias_rows[g] = sum of dY over the group's rows
            M, N = dy.shape
            if M % ctx.rpg != 0:
                raise ValueError("linear_bn_relu: bias_rows needs M % rows_per_group == 0 for the backward")
            drows = dy.view(M // ctx.rpg, ctx.rpg, N).sum(dim=1, dtype=torch.float32)
        return dx, dW, colsum, dgamma, dbeta, None, None, None, None, None, drows, None


def linear_bn_relu(x, W, b, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1, relu=True,
                   bias_rows=None, rows_per_group=0):
    """Fused Linear -> BatchNorm(train) -> ReLU on a (M,K) matrix (bf16 compute).  bias_rows (G,N) adds a
    second bias shared by groups of rows_per_group consecutive rows (a per-object term)."""
    return _LinearBnRelu.apply(x, W, b, gamma, beta, running_mean, running_var, eps, momentum, relu, bias_rows,
                               rows_per_group)


class _MultiLinearBnRelu(torch.autograd.Function):
    """Several Conv1d(k=1) -> BatchNorm1d -> ReLU blocks that read the SAME (M,K) input (the
    1286-channel feature buffer feeds rot_green.conv1, rot_red.conv1, ts.conv1 and
    conv1d_block[0]: reference PoseNet9D.py:37-50, FaceRecon.py:114-116) as one autograd node and
    ONE K6 GEMM each way: the weights are concatenated along the output axis, so the input is read
    from HBM once in the forward, the input gradient is a single dgrad over the concatenated dY
    (accumulated in TMEM instead of read-modify-write passes) and the weight gradients a single wgrad."""

    @staticmethod
    def forward(ctx, x, buffers, *params):
        # params: (W, b, gamma, beta) per block; buffers: [(running_mean, running_var, eps, momentum, relu)]
        xb = _as_gemm_operand(x)
        n = len(buffers)
        Ws = [params[4 * i] for i in range(n)]
        widths = [W.shape[0] for W in Ws]
        tot = sum(widths)
        Wcat = torch.cat(Ws, dim=0).to(torch.bfloat16)                       # (sum N_i, K)
        has_bias = [params[4 * i + 1] is not None for i in range(n)]
        bcat = torch.cat([params[4 * i + 1].float() if has_bias[i] else
                          torch.zeros(widths[i], dtype=torch.float32, device=x.device) for i in range(n)])
        ycat, part = gemm_bf16(xb, Wcat, bias=bcat, stats=True)              # (M, tot), (M/128, 2, tot)
        saved, outs, meta, off = [xb, Wcat, ycat], [], [], 0
        for i in range(n):
            gamma, beta = params[4 * i + 2], params[4 * i + 3]
            rm, rv, eps, mom, relu = buffers[i]
            g32, b32 = gamma.float().contiguous(), beta.float().contiguous()
            z, stats = _bn_fwd_from_partials(ycat[:, off:off + widths[i]], part[:, :, off:off + widths[i]],
                                             g32, b32, rm, rv, eps, mom, relu)
            saved += [g32, b32, stats]
            meta.append((int(relu), has_bias[i], off, widths[i]))
            outs.append(z)
            off += widths[i]
        ctx.save_for_backward(*saved)
        ctx.meta = meta
        ctx.dx_dtype = _dx_dtype(x)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dzs):
        saved = ctx.saved_tensors
        xb, Wcat, ycat = saved[:3]
        M, tot = ycat.shape
        dycat = torch.empty_like(ycat)
        grads = []
        for i, (relu, has_bias, off, wd) in enumerate(ctx.meta):
            g32, b32, stats = saved[3 + 3 * i:6 + 3 * i]
            dz = dzs[i]
            if dz is None:
                dycat[:, off:off + wd].zero_()
                grads.append((None, None, None))
                continue
            _, dgamma, dbeta, colsum = _bn_bwd_raw(ycat[:, off:off + wd], dz, g32, b32, stats, relu, has_bias,
                                                   dx_out=dycat[:, off:off + wd])
            grads.append((colsum, dgamma, dbeta))
        dx = gemm_bf16(dycat, Wcat, b_mn=True, out_dtype=ctx.dx_dtype) if ctx.needs_input_grad[0] else None
        dWcat = _wgrad(dycat, xb)                                            # (tot, K) fp32
        out = []
        for (relu, has_bias, off, wd), (colsum, dgamma, dbeta) in zip(ctx.meta, grads):
            out += [dWcat[off:off + wd] if dgamma is not None else None, colsum, dgamma, dbeta]
        return (dx, None, *out)


def multi_linear_bn_relu(x, blocks):
    """blocks: list of (W, b, gamma, beta, running_mean, running_var, eps, momentum, relu).
    Returns the list of relu?(bn_train(x @ W_i^T + b_i))."""
    params, buffers = [], []
    for W, b, gamma, beta, rm, rv, eps, mom, relu in blocks:
        params += [W, b, gamma, beta]
        buffers.append((rm, rv, eps, mom, relu))
    return list(_MultiLinearBnRelu.apply(x, buffers, *params))


# ------------------------------------------------------------ K6: tensor-core GEMM
def gemm_splits(M, N, K):
    return int(_lib.load().hsp_gemm_bf16_splits(M, N, K, 1))


def gemm_bf16(a, b, a_mn=False, b_mn=False, bias=None, out=None, out_dtype=torch.bfloat16, splits=1,
              stats=False, tile_n=0, ctas=0, bias_rows=None, rows_per_group=0, relu=False, c_in=None):
    """K6: out[M,N] (+bias) = A . B^T on tcgen05 tensor cores (bf16 operands, fp32 accumulate).

    a: (M,K) [a_mn=False] or (K,M) [a_mn=True];  b: (N,K) [b_mn=False] or (K,N) [b_mn=True]; both bf16
    2-D with unit inner stride (row pitch may exceed the width: column slices are fine).
    out_dtype bf16 | fp32; splits > 1 (fp32 only) returns the sum of the split-K planes.
    stats=True also returns the (ceil(M/128), 2, N) BatchNorm partials of the stored values."""
    for t, name in ((a, "a"), (b, "b")):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise _lib.HSPoseLibraryError(f"gemm_bf16: {name} must be a CUDA tensor (no CPU path)")
        if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise TypeError(f"gemm_bf16: {name} must be a 2-D bf16 matrix with unit inner stride")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"gemm_bf16: reduction mismatch {K} vs {Kb}")
    f32 = out_dtype == torch.float32
    with torch.cuda.device(a.device):
        if out is None:
            ldo = N + ((-N) % (4 if f32 else 8))
            buf = torch.empty((splits, M, ldo) if splits > 1 else (M, ldo), dtype=out_dtype, device=a.device)
        else:
            buf, ldo = out, out.stride(-2)
        st = torch.empty((M + 127) // 128, 2, N, dtype=torch.float32, device=a.device) if stats else None
        if bias is not None:
            bias = _need(bias, torch.float32, "bias")
        if bias_rows is not None:
            bias_rows = _need(bias_rows, torch.float32, "bias_rows")
            if bias_rows.shape != ((M + rows_per_group - 1) // rows_per_group, N):
                raise ValueError("gemm_bf16: bias_rows must be (ceil(M / rows_per_group), N)")
        if c_in is not None:      # residual: out = c_in + a . b^T (fp32 output only)
            c_in = _need(c_in, torch.float32, "c_in")
            if c_in.shape != (M, N) or not f32 or splits != 1:
                raise ValueError("gemm_bf16: c_in must be (M, N) fp32 with an fp32, unsplit output")
            _call("hsp_gemm_bf16_acc", _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), M, N, K,
                  _p(bias), _p(bias_rows), int(rows_per_group), int(relu), _p(c_in), c_in.stride(0), _p(buf), ldo,
                  int(f32), splits, _p(st), tile_n, ctas, _stream())
        else:
            _call("hsp_gemm_bf16", _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), M, N, K,
                  _p(bias), _p(bias_rows), int(rows_per_group), int(relu), _p(buf), ldo, int(f32), splits, _p(st),
                  tile_n, ctas, _stream())
    res = buf
    if splits > 1:
        res = buf[0] if splits == 1 else buf.sum(dim=0)
    if out is None and res.shape[-1] != N:
        res = res[..., :N]
    return (res, st) if stats else res


_shadow = None     # (base pointer, bytes, device, bf16 buffer) of the flat fp32 parameter buffer, inside a step


@contextlib.contextmanager
def weight_shadow(flat_fp32, flat_bf16):
    """Within the block, an fp32 GEMM operand that is a view of `flat_fp32` (the engine's flat parameter buffer)
    is served as the same view of `flat_bf16` (its bf16 copy, refreshed by the caller) instead of a cast."""
    global _shadow
    prev = _shadow
    _shadow = (flat_fp32.data_ptr(), flat_fp32.numel() * 4, flat_fp32.device, flat_bf16) if flat_bf16 is not None else None
    try:
        yield
    finally:
        _shadow = prev


def _as_gemm_operand(t):
    """bf16 2-D matrix whose row pitch is 16-byte aligned (TMA requirement); pads columns if needed."""
    if t.dtype == torch.float32 and _shadow is not None and t.device == _shadow[2]:
        off = t.data_ptr() - _shadow[0]
        if 0 <= off < _shadow[1]:
            t = _shadow[3].as_strided(t.shape, t.stride(), off // 4)
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    if t.stride(-1) != 1 or (t.stride(0) * 2) % 16 != 0 or t.data_ptr() % 16 != 0:
        cols = t.shape[1]
        buf = torch.zeros(t.shape[0], cols + ((-cols) % 8), dtype=torch.bfloat16, device=t.device)
        buf[:, :cols] = t
        t = buf[:, :cols]
    return t


def _dx_dtype(x):
    """dtype the input gradient of a tensor-core Linear leaves the GEMM epilogue in: the input's own (fp32 feature
    maps get an fp32 gradient directly — autograd would otherwise cast the bf16 result in a separate pass)."""
    return torch.float32 if x.dtype == torch.float32 else torch.bfloat16


def _wgrad(dy, x):
    """dW (N,K) fp32 = dy^T (M,N) . x (M,K): both operands read MN-major, split-K, planes added in order."""
    M, N = dy.shape
    K = x.shape[1]
    return gemm_bf16(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=gemm_splits(N, K, M))


class _LinearTC(torch.autograd.Function):
    """y (M,N) bf16 = x (M,K) . W (N,K)^T + b on the K6 kernel, forward / dgrad / wgrad (no library GEMM)."""

    @staticmethod
    def forward(ctx, x, W, b):
        xb = _as_gemm_operand(x)
        Wb = _as_gemm_operand(W)
        y = gemm_bf16(xb, Wb, bias=b)
        ctx.save_for_backward(xb, Wb)
        ctx.has_bias = b is not None
        ctx.dx_dtype = _dx_dtype(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, Wb = ctx.saved_tensors
        dyb = _as_gemm_operand(dy)
        dx = gemm_bf16(dyb, Wb, b_mn=True, out_dtype=ctx.dx_dtype) if ctx.needs_input_grad[0] else None
        dW = _wgrad(dyb, xb) if ctx.needs_input_grad[1] else None
        db = dyb.sum(dim=0, dtype=torch.float32) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dW, db


def linear_tc(x, W, b=None):
    """F.linear on the tensor-core kernel: x (..., K), W (N, K) -> (..., N) bf16.  K % 8 == 0."""
    shape = x.shape
    y = _LinearTC.apply(x.reshape(-1, shape[-1]), W, b)
    return y.view(*shape[:-1], W.shape[0])


class _OrlFuse(torch.autograd.Function):
    """The tail of every HS layer on the mixed-precision path as ONE autograd node:
        out = feature + feature @ W2[:, :C]^T + (G @ W2[:, C:]^T)[:, None, :] + STE,   G = get_ORL_global(feature)
    (reference gcn3d.py:109-113 / :183-187 and the `+ f_STE` of :90 / :156; K5a + two K6 GEMMs + K5d).  `feature`
    feeds three consumers; as separate nodes autograd sums their three gradients with two full-size element-wise
    passes over a zero-filled ORL buffer.  Here the pass-through gradient enters the dgrad GEMM as its residual
    (hsp_gemm_bf16_acc) and the ORL backward adds its (order-independent) terms on top: no fill, no add pass."""

    @staticmethod
    def forward(ctx, feature, idx32, W2, ste, xyz, wxyz):
        feature = _need(feature, torch.float32, "feature")
        idx32 = _need(idx32, torch.int32, "idx")
        B, N, C = feature.shape
        k = idx32.shape[2]
        need_grad = ctx.needs_input_grad[0]
        lib = _lib.load()
        W2b = _as_gemm_operand(W2)                                  # (C, 2C) bf16 (a view of the parameter shadow)
        Wf, Wg = W2b[:, :C], W2b[:, C:]
        sdt = F32
        if ste is not None:
            ste = _need(ste, ste.dtype if ste.dtype == torch.bfloat16 else torch.float32, "ste")
            sdt = BF16 if ste.dtype == torch.bfloat16 else F32
        if xyz is not None:
            xyz = _need(xyz, torch.float32, "xyz")
            wxyz = _need(wxyz, torch.float32, "wxyz")
        with torch.cuda.device(feature.device):
            ws = _workspace(lib.hsp_orl_global_workspace_bytes(B, N, C), feature.device)
            G = torch.empty(B, C, dtype=torch.float32, device=feature.device)
            am = torch.empty(B, N, C, dtype=torch.uint8, device=feature.device) if need_grad else None
            _call("hsp_orl_global_fwd", _p(feature), _p(idx32), B, N, C, k, _p(G), _p(am), _p(ws), ws.numel(),
                  _stream())
            f16 = feature.view(B * N, C).to(torch.bfloat16)
            G16 = G.to(torch.bfloat16)
            lin = gemm_bf16(f16, Wf)                                 # (B*N, C) bf16
            gproj = gemm_bf16(G16, Wg).float()                       # (B, C): bf16-rounded like linear_tc(...).float()
            out = torch.empty_like(feature)
            _call("hsp_residual_sum_fwd", _p(feature), _p(lin), BF16, _p(gproj), _p(ste), sdt, _p(xyz), _p(wxyz),
                  B, N, C, _p(out), _stream())
        ctx.save_for_backward(f16, G16, W2b, idx32, am, xyz)
        ctx.meta = (B, N, C, k, None if ste is None else sdt)
        return out

    @staticmethod
    def backward(ctx, g):
        f16, G16, W2b, idx32, am, xyz = ctx.saved_tensors
        B, N, C, k, sdt = ctx.meta
        Wf, Wg = W2b[:, :C], W2b[:, C:]
        g = _need(g.float(), torch.float32, "g")
        want_w = xyz is not None and ctx.needs_input_grad[5]
        with torch.cuda.device(g.device):
            g16 = torch.empty(B, N, C, dtype=torch.bfloat16, device=g.device)
            ggp = torch.empty(B, C, dtype=torch.float32, device=g.device)
            gwp = torch.empty(B, C, 3, dtype=torch.float32, device=g.device) if want_w else None
            _call("hsp_residual_sum_bwd", _p(g), _p(xyz) if want_w else None, B, N, C, _p(g16), _p(ggp), _p(gwp),
                  _stream())
            g16m = g16.view(B * N, C)
            ggp16 = ggp.to(torch.bfloat16)
            dfeat = None
            if ctx.needs_input_grad[0]:
                # d feature = g (pass-through) + g16 @ W2f (the GEMM's residual input) + ORL scatter (on top)
                dfeat = gemm_bf16(g16m, Wf, b_mn=True, out_dtype=torch.float32, c_in=g.view(B * N, C))
                dG = gemm_bf16(ggp16, Wg, b_mn=True, out_dtype=torch.float32)          # (B, C)
                _call("hsp_orl_global_bwd", _p(dG), _p(idx32), _p(am), B, N, C, k, _p(dfeat), _stream())
                dfeat = dfeat.view(B, N, C)
            dW2 = None
            if ctx.needs_input_grad[2]:
                dW2 = torch.cat([_wgrad(g16m, f16), _wgrad(ggp16, G16)], dim=1)         # (C, 2C)
        d_ste = None
        if sdt is not None and ctx.needs_input_grad[3]:
            d_ste = g16 if sdt == BF16 else g
        return dfeat, None, dW2, d_ste, None, (gwp.sum(dim=0) if want_w else None)


def orl_fuse(feature, idx32, W2, ste=None, xyz=None, wxyz=None):
    """Fused ORL + residual (+ STE) tail of an HS layer (mixed-precision path; C % 8 == 0)."""
    return _OrlFuse.apply(feature, idx32, W2, ste, xyz, wxyz)


# ------------------------------------------------------------ K8: fused loss graph
LOSS_TERMS = ("Rot1", "Rot1_cos", "Rot2", "Rot2_cos", "Rot_r_a", "Tran", "Size", "R_con",
              "recon_per_p", "recon_p_f", "recon_point_vote", "recon_point_r", "recon_point_t", "recon_point_s",
              "recon_point_self", "geo_point", "Prop_pm", "Prop_sym_recon", "Prop_sym_rt")
LOSS_WEIGHT_FLAGS = ("rot_1_w", "rot_2_w", "rot_regular", "tran_w", "size_w", "r_con_w", "recon_n_w", "recon_d_w",
                     "recon_f_w", "recon_v_w", "recon_bb_r_w", "recon_bb_t_w", "recon_bb_s_w", "recon_bb_self_w",
                     "geo_p_w", "prop_pm_w", "prop_sym_w")


class _FusedLosses(torch.autograd.Function):
    """All 19 loss terms of stage 'PoseNet_only' (K8): two launches forward, two backward."""

    @staticmethod
    def forward(ctx, weights, face, recon, p_green, p_red, f_green, f_red, pred_T, pred_s, PC, gt_R, gt_t, gt_s,
                mean_shape, sym, obj_id):
        face = _need(face, torch.float32, "face")
        recon = _need(recon, torch.float32, "recon")
        PC = _need(PC, torch.float32, "PC")
        B, N, _ = PC.shape
        if face.shape != (B, N, 30) or recon.shape != (B, N, 3):
            raise ValueError("fused_losses: face (B,N,30), recon (B,N,3), PC (B,N,3)")
        pred = torch.cat([p_green.float(), p_red.float(), f_green.float().view(B, 1), f_red.float().view(B, 1),
                          pred_T.float(), pred_s.float()], dim=1).contiguous()
        gt = torch.cat([gt_R.float().reshape(B, 9), gt_t.float(), gt_s.float(), mean_shape.float(), sym.float(),
                        obj_id.float().view(B, 1)], dim=1).contiguous()
        lib = _lib.load()
        nt, ns = lib.hsp_losses_num_terms(), lib.hsp_losses_num_sums()
        w = (ctypes.c_float * len(weights))(*weights)
        with torch.cuda.device(PC.device):
            sums = torch.empty(B, ns, dtype=torch.float32, device=PC.device)
            pieces = torch.empty(B, nt, dtype=torch.float32, device=PC.device)
            _call("hsp_losses_fwd", _p(face), _p(recon), _p(PC), _p(pred), _p(gt), w, B, N, _p(sums), _p(pieces),
                  _stream())
        ctx.save_for_backward(face, recon, PC, pred, gt, sums)
        ctx.weights = tuple(weights)
        return pieces.sum(dim=0)

    @staticmethod
    def backward(ctx, gterm):
        face, recon, PC, pred, gt, sums = ctx.saved_tensors
        B, N, _ = PC.shape
        gterm = _need(gterm.float(), torch.float32, "gterm")
        w = (ctypes.c_float * len(ctx.weights))(*ctx.weights)
        with torch.cuda.device(PC.device):
            gface = torch.empty_like(face)
            grecon = torch.empty_like(recon)
            gpred = torch.empty_like(pred)
            ws = torch.empty(B, 54, dtype=torch.float32, device=PC.device)
            _call("hsp_losses_bwd", _p(face), _p(recon), _p(PC), _p(pred), _p(gt), w, _p(sums), _p(gterm), B, N,
                  _p(gface), _p(grecon), _p(gpred), _p(ws), _stream())
        return (None, gface, grecon, gpred[:, 0:3], gpred[:, 3:6], gpred[:, 6], gpred[:, 7], gpred[:, 8:11],
                gpred[:, 11:14], None, None, None, None, None, None, None)


def fused_losses(weights, face, recon, p_green, p_red, f_green, f_red, pred_T, pred_s, PC, gt_R, gt_t, gt_s,
                 mean_shape, sym, obj_id):
    """K8 -> dict {term name: scalar} over LOSS_TERMS.  `weights`: the 17 floats of LOSS_WEIGHT_FLAGS."""
    terms = _FusedLosses.apply(tuple(float(x) for x in weights), face, recon, p_green, p_red, f_green, f_red,
                               pred_T, pred_s, PC, gt_R, gt_t, gt_s, mean_shape, sym, obj_id)
    out = LossTerms(zip(LOSS_TERMS, terms.unbind(0)))      # one autograd node for the 19 views
    out.vector = terms
    return out


class LossTerms(dict):
    """{term name: scalar}; `.vector` is the (19,) tensor the scalars are views of (one reduction sums them)."""
    vector = None


class LossGroups(dict):
    """The reference's 4-key loss dict; `.total` (optional) is the sum of every term in it, computed with one
    reduction instead of one add per term — what a train loop that only needs the total should use."""
    total = None


# BatchNorm `num_batches_tracked` of the fused BN paths: one multi-tensor add per forward instead of one launch each
_pending_counters = []


def bump_counter(t):
    if t is not None:
        _pending_counters.append(t)


def flush_counters():
    if _pending_counters:
        torch._foreach_add_(_pending_counters, 1)
        _pending_counters.clear()


# ------------------------------------------------------------ K10: augmentation
def augment(PC, R, t, s, mean_shape, sym, aug_bb, aug_rt_t, aug_rt_r, model_point, nocs_scale, obj_id, gates, ey,
            defor, probs, pc_r):
    """HSPose.data_augment as one launch -> (PC, R, t, s).  gates (B,4), ey (B,2), defor (B,N,3): the uniform
    draws; probs = (aug_bb_pro, aug_rt_pro, aug_bc_pro, aug_pc_pro)."""
    args = [_need(x.float() if x.dtype != torch.float32 else x, torch.float32, n) for x, n in
            ((PC, "PC"), (R, "R"), (t, "t"), (s, "s"), (mean_shape, "mean_shape"), (sym, "sym"), (aug_bb, "aug_bb"),
             (aug_rt_t, "aug_rt_t"), (aug_rt_r, "aug_rt_r"), (model_point, "model_point"), (nocs_scale, "nocs_scale"),
             (obj_id, "obj_id"), (gates, "gates"), (ey, "ey"), (defor, "defor"))]
    B, N, _ = PC.shape
    f = ctypes.c_float
    with torch.cuda.device(PC.device):
        PC_o, R_o = torch.empty_like(args[0]), torch.empty_like(args[1])
        t_o, s_o = torch.empty_like(args[2]), torch.empty_like(args[3])
        _call("hsp_augment", *[_p(a) for a in args], f(probs[0]), f(probs[1]), f(probs[2]), f(probs[3]), f(pc_r),
              B, N, model_point.shape[1], _p(PC_o), _p(R_o), _p(t_o), _p(s_o), _stream())
    return PC_o, R_o, t_o, s_o


# ------------------------------------------------------------ K11: input pre-stage (depth ROI -> sampled cloud)
def depth_to_cloud(depth, mask, xymap, camK):
    """Back-projection + order-preserving compaction of the valid pixels (K11).
    depth, mask (B,H,W) or (B,1,H,W); xymap (B,2,H,W); camK (B,3,3) — float64 selects the numpy (float64)
    arithmetic of datasets/load_data.py:322-333, float32 the torch arithmetic of pc_sample.py:24-54.
    -> cloud (B, H*W, 3) fp32 metres (rows >= count[b] unspecified), count (B,) int32."""
    if depth.dim() == 4:
        depth = depth[:, 0]
    if mask.dim() == 4:
        mask = mask[:, 0]
    depth = _need(depth.float(), torch.float32, "depth")
    mask = _need(mask.float(), torch.float32, "mask")
    xymap = _need(xymap.float(), torch.float32, "xymap")
    B, H, W = depth.shape
    if mask.shape != (B, H, W) or xymap.shape != (B, 2, H, W) or camK.shape != (B, 3, 3):
        raise ValueError("depth_to_cloud: depth/mask (B,H,W), xymap (B,2,H,W), camK (B,3,3)")
    f64 = camK.dtype == torch.float64
    camK = _need(camK, torch.float64 if f64 else torch.float32, "camK")
    with torch.cuda.device(depth.device):
        cloud = torch.empty(B, H * W, 3, dtype=torch.float32, device=depth.device)
        count = torch.empty(B, dtype=torch.int32, device=depth.device)
        _call("hsp_depth_to_cloud", _p(depth), _p(mask), _p(xymap), _p(camK), int(f64), B, H, W, _p(cloud),
              _p(count), _stream())
    return cloud, count


def sample_points(cloud, count, n_pts, choose=None, seed=0, status=None):
    """n_pts points per object from the compacted cloud (K11).  choose (B,n_pts) int: the caller's draw
    (gather); None: the device rule — tile when count <= n_pts (load_data.py:316-317), else a random subset
    without replacement keyed by `seed`.  status: optional int32 device scalar (see the header)."""
    cloud = _need(cloud, torch.float32, "cloud")
    count = _need(count, torch.int32, "count")
    B, cap, _ = cloud.shape
    if choose is not None:
        choose = _need(choose.to(torch.int32), torch.int32, "choose")
        if choose.shape != (B, n_pts):
            raise ValueError("sample_points: choose must be (B, n_pts)")
    with torch.cuda.device(cloud.device):
        out = torch.empty(B, n_pts, 3, dtype=torch.float32, device=cloud.device)
        _call("hsp_sample_points", _p(cloud), _p(count), _p(choose), ctypes.c_ulonglong(int(seed) & (2 ** 64 - 1)),
              B, cap, int(n_pts), _p(out), _p(status), _stream())
    return out


# ------------------------------------------------------------ fp32-accurate GEMM on the bf16 tensor cores
# term order: SMALLEST first (x1 w3, x2 w2, x3 w1 ~ 2^-16; x1 w2, x2 w1 ~ 2^-8; x1 w1 last), so the small terms are
# summed while the TMEM accumulator is still small — with the large term first a single accumulator loses them
# to its rounding (measured 9e-6 vs 9e-7 relative error, tools/split_check.py)
_SPLIT_A = (0, 1, 2, 0, 1, 0)
_SPLIT_B = (2, 1, 0, 1, 0, 0)


def split_bf16(x, comp, kpad):
    """(M,K) fp32 -> (M, len(comp)*kpad) bf16 concatenation of the bf16 split components comp[t] of x."""
    x = _need(x, torch.float32, "x")
    if x.dim() != 2:
        raise ValueError("split_bf16 expects a 2-D matrix")
    M, K = x.shape
    arr = (ctypes.c_int * len(comp))(*comp)
    with torch.cuda.device(x.device):
        out = torch.empty(M, len(comp) * kpad, dtype=torch.bfloat16, device=x.device)
        _call("hsp_split_bf16", _p(x), x.stride(0), M, K, kpad, len(comp), arr, _p(out), _stream())
    return out


def linear_fp32x(x, W, b=None, relu=False, w_split=None, splits=1):
    """y = x @ W^T (+ b) (ReLU) with fp32 accuracy on the bf16 tensor cores: 3-way bf16 split of both operands,
    the six significant cross terms as ONE K6 GEMM over a 6x longer reduction axis (measured on B200: error vs
    float64 on par with the strict-fp32 library GEMM, tools/split_check.py; splits = 6 gives every term its own
    accumulator plane).  x (M,K), W (N,K) fp32 -> (M,N) fp32.  No autograd (evaluation path).
    w_split: the split of W from a previous call (weights are constant in evaluation)."""
    M, K = x.shape
    kpad = (K + 63) // 64 * 64
    a = split_bf16(x, _SPLIT_A, kpad)
    bw = w_split if w_split is not None else split_bf16(W, _SPLIT_B, kpad)
    if splits == 1:      # bias and ReLU in the GEMM epilogue
        return gemm_bf16(a, bw, bias=b, relu=relu, out_dtype=torch.float32)
    y = gemm_bf16(a, bw, out_dtype=torch.float32, splits=splits)
    if b is not None:
        y = y + b
    return torch.relu_(y) if relu else y


class _EvalLinear(torch.autograd.Function):
    """linear_fp32x as an autograd node that refuses to back-propagate (evaluation fast path: the folded,
    pre-split weights carry no gradient) — loud instead of a silently missing gradient."""

    @staticmethod
    def forward(ctx, x, w_split, b, relu, n_out):
        M, K = x.shape
        kpad = w_split.shape[1] // 6
        a = split_bf16(x, _SPLIT_A, kpad)
        return gemm_bf16(a, w_split, bias=b, relu=relu, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("hs-pose_b200: the folded Conv1d+BatchNorm(eval) fast path has no backward; "
                                  "call .train() (batch statistics) to differentiate through this block")


def eval_linear(x, w_split, b, relu):
    return _EvalLinear.apply(x, w_split, b, relu, w_split.shape[0])


# ---------------------------------------------------------------- chamfer
class _Chamfer(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, a, b):
        a = _need(a, torch.float32, "a")
        b = _need(b, torch.float32, "b")
        B, N, _ = a.shape
        M = b.shape[1]
        with torch.cuda.device(a.device):
            da = torch.empty(B, N, dtype=torch.float32, device=a.device)
            db = torch.empty(B, M, dtype=torch.float32, device=a.device)
            ia = torch.empty(B, N, dtype=torch.int32, device=a.device)
            ib = torch.empty(B, M, dtype=torch.int32, device=a.device)
            _call("hsp_chamfer_fwd", _p(a), _p(b), B, N, M, _p(da), _p(ia), _p(db), _p(ib),
                  _stream())
        ctx.save_for_backward(a, b, ia, ib)
        ctx.mark_non_differentiable(ia, ib)
        return da, db, ia, ib

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gda, gdb, _gia, _gib):
        a, b, ia, ib = ctx.saved_tensors
        B, N, _ = a.shape
        M = b.shape[1]
        gda = _need(gda.float(), torch.float32, "gda")
        gdb = _need(gdb.float(), torch.float32, "gdb")
        with torch.cuda.device(a.device):
            ga = torch.empty_like(a)
            gb = torch.empty_like(b)
            _call("hsp_chamfer_bwd", _p(a), _p(b), _p(ia), _p(ib), _p(gda), _p(gdb), B, N, M,
                  _p(ga), _p(gb), _stream())
        return ga, gb


def chamfer(a, b):
    """K7: (dist_a (B,N), dist_b (B,M), idx_a, idx_b) squared NN distances both ways."""
    return _Chamfer.apply(a, b)
