"""K6 check on the GPU: hsp_gemm_bf16 vs torch fp32 matmul of the same bf16 operands; optional timing.
   python tools/gemm_check.py case M N K a_mn b_mn f32 splits tile_n ctas [time]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hspose_b200.ops as ops


def run(M, N, K, a_mn, b_mn, f32, splits, tile_n, ctas, timeit):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev) if splits == 1 else None
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    if bias is not None:
        ref = ref + bias
    want_stats = (not f32) and splits == 1
    out = ops.gemm_bf16(a, b, a_mn, b_mn, bias=bias, out_dtype=torch.float32 if f32 else torch.bfloat16,
                        splits=splits, stats=want_stats, tile_n=tile_n, ctas=ctas)
    if want_stats:
        out, st = out
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-2 if not f32 else 2e-3) * scale
    msg = f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} f32={f32} splits={splits} tile_n={tile_n} ctas={ctas}: max_err={err:.4g} (scale {scale:.3g})"
    ok = err <= tol
    if want_stats:
        y = out.float()
        s_ref, q_ref = y.sum(0), (y * y).sum(0)
        s, q = st[:, 0].sum(0), st[:, 1].sum(0)
        e1 = ((s - s_ref).abs() / (s_ref.abs() + 1e-3 * M ** 0.5)).max().item()
        e2 = ((q - q_ref).abs() / (q_ref.abs() + 1e-6)).max().item()
        msg += f" stats_rel_err=({e1:.3g},{e2:.3g})"
        ok = ok and e1 < 1e-2 and e2 < 1e-3
    if timeit:
        for _ in range(3):
            ops.gemm_bf16(a, b, a_mn, b_mn, bias=bias, out_dtype=torch.float32 if f32 else torch.bfloat16,
                          splits=splits, stats=want_stats, tile_n=tile_n, ctas=ctas)
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.gemm_bf16(a, b, a_mn, b_mn, bias=bias, out_dtype=torch.float32 if f32 else torch.bfloat16,
                          splits=splits, stats=want_stats, tile_n=tile_n, ctas=ctas)
        e1_.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1_) / reps
        # library baseline
        Ab, Bb = A, B
        for _ in range(3):
            torch.matmul(Ab, Bb.t())
        e0.record()
        for _ in range(reps):
            torch.matmul(Ab, Bb.t())
        e1_.record()
        torch.cuda.synchronize()
        ms_lib = e0.elapsed_time(e1_) / reps
        msg += f" | {ms:.4f} ms = {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s (cuBLAS NT {ms_lib:.4f} ms = {2.0 * M * N * K / ms_lib / 1e9:.0f})"
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    return ok


if __name__ == "__main__":
    a = sys.argv[1:]
    M, N, K, a_mn, b_mn, f32, splits, tile_n, ctas = [int(x) for x in a[:9]]
    ok = run(M, N, K, bool(a_mn), bool(b_mn), bool(f32), splits, tile_n, ctas, len(a) > 9)
    sys.exit(0 if ok else 1)
