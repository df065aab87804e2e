#!/usr/bin/env python
"""K2-TC timing (feature-space KNN: split + norm + tensor-core filter + exact refine) on the step's shapes;
HSP_KNN_FEAT_EXACT=1 times the all-FP32 kernel instead.  (The round-1 filter kernel this was A/B-ed against is in
git history before "K2-TC v2"; numbers in profiles/r2_k2_ab.md.)  Output: one JSON line per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.ops as ops  # noqa: E402
from oracle import c_oracle as co  # noqa: E402  (checker only)

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


tag = "fp32-exact" if os.environ.get("HSP_KNN_FEAT_EXACT") else "tc"
for (B, N, D, k) in ((128, 1028, 128, 20), (128, 257, 128, 20), (128, 257, 256, 20), (16, 1028, 256, 20), (8, 3000, 128, 32)):
    g = torch.Generator().manual_seed(N + D)
    f = torch.relu(torch.randn(B, N, D, generator=g) + 0.5).to(dev)
    try:
        got = ops.knn_feat(f, k, want64=True)[0]
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(kernel=tag, B=B, N=N, D=D, k=k, error=str(e)[:80])), flush=True)
        continue
    exact = bool((got[:2].cpu().numpy() == co.neighbor_index(f[:2].cpu().numpy(), k)).all())
    ms = timeit(lambda: ops.knn_feat(f, k))
    print(json.dumps(dict(kernel=tag, B=B, N=N, D=D, k=k, ms=ms, bit_exact_vs_oracle_first_2_objects=exact)), flush=True)
