#!/usr/bin/env python
"""BASELINE.json configs[1]: fp32 eval forward of the drop-in PoseNet9D (FLAGS.train = 0, eval mode — the
path evaluation/evaluate.py:91-98 drives, batch = detections of one image) on one B200.

Reports ms per forward and objects/s for B in {1, 4, 8, 16}, eagerly launched and replayed from a CUDA
graph (the reference's published metric is 38 images/s on an RTX 3090 at ~1-8 objects per image,
supplementary Table 1).  Output: JSON lines on stdout and gpurun_out/eval_bench.jsonl.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hspose_b200.flags as hf  # noqa: E402
from hspose_b200 import gcn3d  # noqa: E402
from hspose_b200.PoseNet9D import PoseNet9D  # noqa: E402
from hspose_b200.synth import fill_params, synth_batch  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
F = hf.get_flags()
F.train, F.gcn_n_num = 0, 20
net = fill_params(PoseNet9D()).to(dev).eval()


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []
for B in (1, 4, 8, 16, 64):
    batch = synth_batch(B, 1028, seed=1, train=False)
    pc, oid = batch["PC"].to(dev), batch["obj_id"].to(dev)
    rows32 = [torch.randperm(1028)[:257].to(dev, torch.int32), torch.randperm(257)[:64].to(dev, torch.int32)]
    it = iter(())

    def provider(vn, pn, device, _state={"i": 0}):
        r = rows32[_state["i"] % 2]
        _state["i"] += 1
        return r

    def fwd():
        with torch.no_grad():
            return net(pc, oid)

    prev = gcn3d.set_pool_rows_provider(provider)   # fixed pooling rows: the forward is graph-capturable
    try:
        eager = timed(fwd)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fwd()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            out = fwd()
        graph = timed(g.replay)
    finally:
        gcn3d.set_pool_rows_provider(prev)
    row = dict(config="eval forward fp32, N=1028, k=20", B=B, ms_eager=eager, ms_graph=graph,
               objects_per_s_eager=B / eager * 1e3, objects_per_s_graph=B / graph * 1e3)
    rows.append(row)
    print(json.dumps(row), flush=True)
# the evaluation loop's per-image call through engine.EvalRunner (forward + generate_RT, bucketed CUDA graphs)
from hspose_b200.engine import EvalRunner  # noqa: E402
from hspose_b200.HSPose import HSPose  # noqa: E402
model = HSPose("PoseNet_only").to(dev).eval()
model.posenet = net
runner = EvalRunner(model)
for B in (1, 2, 3, 5, 8, 16):
    b = synth_batch(B, 1028, seed=1, train=False)
    args = [b[k].to(dev) for k in ("PC", "obj_id", "mean_shape", "sym")]
    ms = timed(lambda: runner(*args), reps=50)
    row = dict(config="EvalRunner: eval forward fp32 + generate_RT, CUDA graph per bucket, N=1028, k=20", B=B, ms=ms,
               images_per_s=1e3 / ms, objects_per_s=B / ms * 1e3)
    rows.append(row)
    print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/eval_bench.jsonl", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
