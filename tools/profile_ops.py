#!/usr/bin/env python
"""torch.profiler view of the bench train step grouped by aten op with input shapes (which torch-native
kernels are worth fusing).  Output: gpurun_out/profile_ops.txt"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hspose_b200.engine import TrainStep
from hspose_b200.HSPose import HSPose
from hspose_b200.synth import synth_batch
from torch.profiler import ProfilerActivity, profile
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = HSPose("PoseNet_only", chamfer_w=1.0).to(dev).train()
tr = TrainStep(model, lr=1e-4, clip=5.0, amp=True, graph=False)
batch = {k: v.to(dev) for k, v in synth_batch(B, 1028, seed=1, train=True).items()}
for _ in range(3):
    tr(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    tr(batch)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/profile_ops.txt", "w") as f:
    f.write(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=120,
                                                                max_name_column_width=60, max_shapes_column_width=90))
    f.write("\n\n")
    f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=80, max_name_column_width=80))

    f.write("\n\n==== aten ops by python call site (self CUDA time) ====\n")
    rows = [e for e in prof.key_averages(group_by_stack_n=8) if e.key.startswith("aten::") and e.self_device_time_total > 0]
    rows.sort(key=lambda e: -e.self_device_time_total)
    tot = sum(e.self_device_time_total for e in rows)
    f.write(f"total self CUDA time of aten ops: {tot / 1e3:.3f} ms\n")
    for e in rows[:120]:
        site = [s for s in e.stack if "hs-pose_b200" in s or "hspose_b200" in s or "bench.py" in s][:3]
        f.write(f"{e.self_device_time_total:9.1f} us  x{e.count:<4d} {e.key:34s} {' <- '.join(x.split('/')[-1] for x in site)}\n")

    f.write("\n\n==== aten ops by input shape (self CUDA time) ====\n")
    rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
    rows.sort(key=lambda e: -e.self_device_time_total)
    for e in rows[:150]:
        f.write(f"{e.self_device_time_total:9.1f} us  x{e.count:<4d} {e.key:34s} {e.input_shapes}\n")
