#!/usr/bin/env python
"""Per-kernel micro-benchmark (CUDA events, L2 flushed between reps)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = []
    for B in (16, 128):
        g = torch.Generator().manual_seed(0)
        N, k, S = 1028, 20, 7
        xyz = (torch.randn(B, N, 3, generator=g) * 0.05).to(dev)
        ms = timeit(lambda: ops.knn3(xyz, xyz, k))
        res.append(dict(kernel="knn3", B=B, N=N, k=k, ms=ms, alg_GBs=B * (12 * N + 4 * N * k) / ms / 1e6))
        idx = ops.knn3(xyz, xyz, k)[1]
        for C, Cin in ((128, 128),):
            fm = torch.relu(torch.randn(B, N, Cin, generator=g)).to(dev)
            ms = timeit(lambda: ops.knn_feat(fm, k))
            res.append(dict(kernel="knn_feat", B=B, N=N, D=Cin, k=k, ms=ms,
                            tflops=2.0 * B * N * N * Cin / ms / 1e9))
            rf = ops.knn_feat(fm, k)[1]
            dirn = torch.nn.functional.normalize(torch.randn(3, S * C, generator=g), dim=0).to(dev)
            P = torch.randn(B, N, (S + 1) * C, generator=g).to(dev)
            ms = timeit(lambda: ops.surface_conv(xyz, idx, dirn, S, C))
            res.append(dict(kernel="surface_conv_fwd", B=B, N=N, C=C, ms=ms))
            ms = timeit(lambda: ops.graph_conv(xyz, rf, dirn, P, S, C))
            res.append(dict(kernel="graph_conv_fwd_f32_nograd", B=B, N=N, C=C, ms=ms,
                            gather_GBs=B * N * k * S * C * 4 / ms / 1e6))
            ms = timeit(lambda: ops._graph_conv_fwd_raw(xyz, rf, dirn, P, S, C, True))
            res.append(dict(kernel="graph_conv_fwd_f32_argmax", B=B, N=N, C=C, ms=ms))
            P16 = P.to(torch.bfloat16)
            ms = timeit(lambda: ops._graph_conv_fwd_raw(xyz, rf, dirn, P16, S, C, True))
            res.append(dict(kernel="graph_conv_fwd_bf16_tagged", B=B, N=N, C=C, ms=ms,
                            gather_GBs=B * N * k * S * C * 2 / ms / 1e6))
            ms = timeit(lambda: ops._graph_conv_fwd_raw(xyz, rf, dirn, P16, S, C, False))
            res.append(dict(kernel="graph_conv_fwd_bf16_nograd", B=B, N=N, C=C, ms=ms))
            feat = torch.randn(B, N, C, generator=g).to(dev)
            ms = timeit(lambda: ops.orl_global(feat, idx))
            res.append(dict(kernel="orl_global", B=B, N=N, C=C, ms=ms))
            Pg = P.clone().requires_grad_()
            dg = dirn.clone().requires_grad_()
            out = ops.graph_conv(xyz, rf, dg, Pg, S, C)
            go = torch.randn_like(out)
            ms = timeit(lambda: torch.autograd.grad(out, (Pg, dg), go, retain_graph=True))
            res.append(dict(kernel="graph_conv_bwd", B=B, N=N, C=C, ms=ms))
    for r in res:
        print(json.dumps(r))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench.jsonl", "w") as f:
        for r in res:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
