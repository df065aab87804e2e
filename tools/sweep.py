#!/usr/bin/env python
"""BASELINE.json configs[4]: N-sweep {256,512,1028,2048,4096} x k {8,16,32} of the KNN + 3D-GCN layer
kernels on one B200 — achieved algorithmic GB/s against the measured HBM roofline.

Layer shapes: HSlayer_surface(128, 7) ("conv_0") and HS_layer(128, 128, 7) ("conv_1"); B = ceil(128*1028/N)
objects; fp32 P.  Timing: CUDA events, L2 flushed between repetitions, median of 7.  Output: one JSON line
per (shape, N, k) to stdout and gpurun_out/sweep.jsonl.
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops  # noqa: E402
from kbench import timeit  # noqa: E402

dev = torch.device("cuda:0")
S, C = 7, 128
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def main():
    out = []
    g = torch.Generator().manual_seed(0)
    for N in (256, 512, 1028, 2048, 4096):
        B = math.ceil(128 * 1028 / N)
        xyz = (torch.randn(B, N, 3, generator=g) * 0.05).to(dev)
        fm = torch.relu(torch.randn(B, N, C, generator=g)).to(dev)
        dirn = torch.nn.functional.normalize(torch.randn(3, S * C, generator=g), dim=0).to(dev)
        W = (torch.randn(C, (S + 1) * C, generator=g) / C ** 0.5).to(dev)
        for k in (8, 16, 32):
            t_knn3 = timeit(lambda: ops.knn3(xyz, xyz, k), reps=7)
            idx = ops.knn3(xyz, xyz, k)[1]
            t_knnf = timeit(lambda: ops.knn_feat(fm, k), reps=7)
            rf = ops.knn_feat(fm, k)[1]
            t_surf = timeit(lambda: ops.surface_conv(xyz, idx, dirn, S, C), reps=7)
            P = (fm.view(-1, C) @ W).view(B, N, (S + 1) * C)
            t_gemm = timeit(lambda: fm.view(-1, C) @ W, reps=7)
            t_gc = timeit(lambda: ops.graph_conv(xyz, rf, dirn, P, S, C), reps=7)
            feat = ops.graph_conv(xyz, rf, dirn, P, S, C)
            t_orl = timeit(lambda: ops.orl_global(feat, idx), reps=7)
            # algorithmic bytes per object: stand-alone KNN and whole layers
            knn3_b = 4 * N * 3 + 8 * N * k
            knnf_b = 4 * N * C + 8 * N * k
            l0_b = 4 * N * 3 + 2 * 2 * 4 * N * k + 3 * 4 * N * C                      # surface layer
            l1_b = 4 * N * (3 + C) + 2 * 4 * N * (S + 1) * C + 2 * 2 * 4 * N * k + 3 * 4 * N * C
            l0_ms = t_knn3 + t_surf + t_orl
            l1_ms = t_knn3 + t_knnf + t_gemm + t_gc + t_orl
            row = dict(N=N, k=k, B=B, ms=dict(knn3=t_knn3, knn_feat=t_knnf, surface_conv=t_surf, gemm_P=t_gemm,
                                              graph_conv=t_gc, orl=t_orl),
                       GBps=dict(knn3=B * knn3_b / t_knn3 / 1e6, knn_feat=B * knnf_b / t_knnf / 1e6,
                                 layer_surface=B * l0_b / l0_ms / 1e6, layer_hs=B * l1_b / l1_ms / 1e6),
                       frac_of_hbm=dict(layer_surface=B * l0_b / l0_ms / 1e6 / PEAK,
                                        layer_hs=B * l1_b / l1_ms / 1e6 / PEAK),
                       objects_per_s=dict(layer_surface=B / l0_ms * 1e3, layer_hs=B / l1_ms * 1e3),
                       alg_bytes_per_object=dict(knn3=knn3_b, knn_feat=knnf_b, layer_surface=l0_b, layer_hs=l1_b),
                       peak_GBps=PEAK)
            out.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep.jsonl", "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
