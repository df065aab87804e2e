#!/usr/bin/env python
"""Where do the remaining library (ATen) launches of the train step come from?

Runs the engine's eager step body under torch.profiler with Python stacks and attributes every CUDA kernel
that is not one of ours to (a) the innermost hs-pose_b200 source line that issued it (forward) or (b) the
autograd node that issued it (backward).  Output: gpurun_out/glue_profile.txt
"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hspose_b200 import engine  # noqa: E402
from hspose_b200.HSPose import HSPose  # noqa: E402
from hspose_b200.synth import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = HSPose("PoseNet_only", chamfer_w=1.0).to(dev).train()
step = engine.TrainStep(model, amp=True, graph=False)
batch = {k: v.to(dev) for k, v in synth_batch(B, 1028, seed=1, train=True).items()}
for _ in range(3):
    step(batch)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    step(batch)
    torch.cuda.synchronize()

OWN = ("hsp::",)


def short(k):
    k = re.sub(r"\(.*$", "", k).replace("void ", "")
    m = re.match(r"at::native::(\w+)<.*?(\w+(?:Functor|_kernel_cuda|kernel_impl|Ops|functor)\w*)", k)
    return (m.group(1)[:12] + ":" + m.group(2)) if m else k[:60]


def site(evt):
    for fr in evt.stack or []:
        if "hs-pose_b200" in fr or "hspose_b200" in fr:
            m = re.search(r"(hs-pose_b200|hspose_b200)/(\S+)\((\d+)\): (\w+)", fr)
            if m:
                return f"{m.group(2)}:{m.group(3)} {m.group(4)}"
    e = evt
    top = None
    while e is not None:
        if e.name.startswith("autograd::engine::evaluate_function"):
            top = e.name.split(": ", 1)[-1]
        e = e.cpu_parent
    return "bwd " + (top or "?")


agg = collections.defaultdict(lambda: [0, 0.0])
seen = set()
for evt in prof.events():
    if not evt.kernels:
        continue
    # only leaf CPU ops own their kernels uniquely; dedupe by kernel identity
    for k in evt.kernels:
        key = (k.name, id(k))
        if key in seen:
            continue
        if any(o in k.name for o in OWN):
            continue
        # attribute to the deepest event that lists the kernel: events() is in start order, children later,
        # so remember and overwrite
        agg_key = (site(evt), short(k.name))
        seen.add(key)
        agg[agg_key][0] += 1
        agg[agg_key][1] += k.duration
lines = [f"# B={B}: library kernels of ONE eager train step by issuing source line / autograd node",
         f"{'count':>5s} {'us':>8s}  site | kernel"]
tot_n = tot_t = 0
for (s, k), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{n:5d} {t:8.1f}  {s} | {k}")
    tot_n += n
    tot_t += t
lines.append(f"total {tot_n} launches, {tot_t / 1e3:.3f} ms")
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/glue_profile.txt", "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
