#!/usr/bin/env python
"""A/B of the two K4b (graph_conv backward) kernels — global float atomics (+ cast pass for a bf16 gP) vs
object-resident shared-memory slabs — on the three HS-layer shapes of the B=128 step.
Eager launches timed with CUDA events, L2 flushed between repetitions; gradients compared with the atomic kernel.  Output: gpurun_out/k4b_variants.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops  # noqa: E402
from hspose_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
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


rows = []
S, k = 7, 20
variants = [1]
for (B, N, C, dt) in ((128, 1028, 128, torch.bfloat16), (128, 257, 256, torch.bfloat16), (128, 64, 512, torch.bfloat16),
                      (128, 1028, 128, torch.float32)):
    g = torch.Generator().manual_seed(0)
    xyz = (torch.randn(B, N, 3, generator=g) * 0.05).to(dev)
    kk = min(k, N // 8) if N < 1028 else k
    idx = ops.knn3(xyz, xyz, kk)[1]
    dirn = torch.nn.functional.normalize(torch.randn(3, S * C, generator=g), dim=0).to(dev)
    P = torch.randn(B, N, (S + 1) * C, generator=g).to(dev).to(dt)
    out, am = ops._graph_conv_fwd_raw(xyz, idx, dirn, P, S, C, True)
    gout = torch.randn(B, N, C, generator=g).to(dev)
    ref = None
    for v in variants:
        run = lambda: ops._graph_conv_bwd_raw(xyz, idx, dirn, P, am, gout, S, C, want_gbias=True, variant="atomic")  # noqa: E731
        res = run()
        if ref is None:
            ref = res
        err = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(res, ref)]
        ms = timeit(run)
        run16 = lambda: ops._graph_conv_bwd_raw(xyz, idx, dirn, P, am, gout, S, C, want_gbias=True, variant="atomic",  # noqa: E731
                                                gp_dtype=torch.bfloat16)
        ms16 = timeit(run16)
        rows.append(dict(B=B, N=N, C=C, k=kk, P=str(dt).split(".")[-1], variant=f"atomic v{v}", ms_fp32_gP=ms,
                         ms_bf16_gP_incl_cast=ms16, rel_err_gP_gdirn_gbias=err))
        print(json.dumps(rows[-1]), flush=True)
    for gdt in (torch.float32, torch.bfloat16):
        run = lambda: ops._graph_conv_bwd_raw(xyz, idx, dirn, P, am, gout, S, C, want_gbias=True, variant="obj",  # noqa: E731
                                              gp_dtype=gdt)
        res = run()
        err = [float((a.float() - b).abs().max() / b.abs().max()) for a, b in zip(res, ref)]
        again = run()
        same = all(bool(torch.equal(a, b)) for a, b in zip(res, again))
        ms = timeit(run)
        rows.append(dict(B=B, N=N, C=C, k=kk, P=str(dt).split(".")[-1], bit_reproducible=same, variant="obj", gP=str(gdt).split(".")[-1], ms=ms,
                         rel_err_gP_gdirn_gbias=err))
        print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/k4b_variants.jsonl", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
