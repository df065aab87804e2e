#!/usr/bin/env python
"""Experiment: K4b (graph_conv backward, fp32 atomics into gP) run per chunk of objects with an L2-sized fp32
scratch that is cast to bf16 right away, against the whole-batch launch + separate cast pass.  CUDA-graph
replays, L2 flushed between replays.  Output: gpurun_out/k4b_chunk.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8, warm=2):
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


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


rows = []
S, k = 7, 20
for (B, N, C) in ((128, 1028, 128), (128, 257, 256), (128, 64, 512)):
    g = torch.Generator().manual_seed(0)
    xyz = (torch.randn(B, N, 3, generator=g) * 0.05).to(dev)
    kk = min(k, N // 8) if N < 1028 else k
    idx = ops.knn3(xyz, xyz, kk)[1]
    dirn = torch.nn.functional.normalize(torch.randn(3, S * C, generator=g), dim=0).to(dev)
    P = torch.randn(B, N, (S + 1) * C, generator=g).to(dev).to(torch.bfloat16)
    out, am = ops._graph_conv_fwd_raw(xyz, idx, dirn, P, S, C, True)
    gout = torch.randn(B, N, C, generator=g).to(dev)
    LD = (S + 1) * C
    gP16 = torch.empty(B, N, LD, dtype=torch.bfloat16, device=dev)

    def whole():
        gP, gd, gb = ops._graph_conv_bwd_raw(xyz, idx, dirn, P, am, gout, S, C, want_gbias=True)
        gP16.copy_(gP)

    def only_kernel():
        ops._graph_conv_bwd_raw(xyz, idx, dirn, P, am, gout, S, C, want_gbias=True)

    ref = None
    for name, fn in (("whole+cast", whole), ("whole kernel only (memset + K4b)", only_kernel)):
        ms = timeit(graphed(fn))
        rows.append(dict(B=B, N=N, C=C, variant=name, ms=ms))
        print(json.dumps(rows[-1]), flush=True)
    ref = gP16.clone()
    for chunk in (4, 8, 12, 16, 24, 32, 64):
        def chunked():
            for b0 in range(0, B, chunk):
                b1 = min(B, b0 + chunk)
                gP, gd, gb = ops._graph_conv_bwd_raw(xyz[b0:b1], idx[b0:b1], dirn, P[b0:b1], am[b0:b1], gout[b0:b1],
                                                     S, C, want_gbias=True)
                gP16[b0:b1].copy_(gP)
        ms = timeit(graphed(chunked))
        err = (gP16.float() - ref.float()).abs().max().item()
        rows.append(dict(B=B, N=N, C=C, variant=f"chunks of {chunk} objects ({chunk * N * LD * 4 / 2**20:.0f} MiB scratch)",
                         ms=ms, max_abs_diff_vs_whole=err))
        print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/k4b_chunk.jsonl", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
