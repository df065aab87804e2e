"""Accuracy of the 6-term bf16 split GEMM (ops.linear_fp32x) vs float64, next to torch's strict fp32 GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hspose_b200.ops as ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
for (M, N, K) in [(16448, 1024, 1296), (16448, 256, 1024), (16448, 1024, 128), (4112, 2048, 256)]:
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.relu(torch.randn(M, K, generator=g)).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    ref = (x.double() @ W.double().t() + b.double())
    y32 = torch.addmm(b, x, W.t())
    ys = ops.linear_fp32x(x, W, b, splits=int(os.environ.get("SPLITS", "1")))
    sc = ref.abs().max().item()
    e32 = (y32.double() - ref).abs().max().item() / sc
    es = (ys.double() - ref).abs().max().item() / sc
    r32 = ((y32.double() - ref).norm() / ref.norm()).item()
    rs = ((ys.double() - ref).norm() / ref.norm()).item()
    def t(fn):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    wsp = ops.split_bf16(W, ops._SPLIT_B, (K + 63) // 64 * 64)
    print(f"M={M} N={N} K={K}: max err/scale fp32 {e32:.2e} split {es:.2e} | rel l2 fp32 {r32:.2e} split {rs:.2e} | "
          f"ms fp32 {t(lambda: torch.addmm(b, x, W.t())):.3f} split {t(lambda: ops.linear_fp32x(x, W, b, w_split=wsp, splits=int(os.environ.get('SPLITS', '1')))):.3f}", flush=True)
