import os, sys, ctypes, torch
sys.path.insert(0, "/root/repo")
import hspose_b200.ops as ops
from hspose_b200 import _lib, gcn3d
from hspose_b200.HSPose import HSPose
from hspose_b200.synth import synth_batch
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 16
model = HSPose("PoseNet_only").to(dev).train()
batch = synth_batch(B, 1028, seed=1, train=True)
pc = batch["PC"].to(dev)
v = (pc - pc.mean(1, keepdim=True)).contiguous()
fr = model.posenet.face_recon
with torch.no_grad():
    fm0 = torch.relu(fr.conv_0(v, 20)).contiguous()
print("fm0 stats: mean |f|^2", (fm0 ** 2).sum(-1).mean().item(), "max", (fm0 ** 2).sum(-1).max().item(), "zeros frac", (fm0 == 0).float().mean().item())
d = torch.cdist(fm0[0], fm0[0]) ** 2
ds, _ = d.sort(dim=1)
print("d2 NN mean", ds[:, 1].mean().item(), "20th", ds[:, 20].mean().item(), "gap20-21 median", (ds[:, 21] - ds[:, 20]).median().item(), "count within +0.05 of 20th:", ((d <= ds[:, 20:21] + 0.05).sum(1).float().mean().item()))
lib = _lib.load()
N, D, k = 1028, 128, 20
ws = torch.zeros(lib.hsp_knn_feat_workspace_bytes(B, N), dtype=torch.uint8, device=dev)
i32 = torch.empty(B, N, k, dtype=torch.int32, device=dev)
rc = lib.hsp_knn_feat(ctypes.c_void_p(fm0.data_ptr()), B, N, D, k, 1, None, ctypes.c_void_p(i32.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
T64 = 2 * ((N + 127) // 128)
base = (ws.data_ptr() + 127) & ~127
off = base - ws.data_ptr()
off += B * T64 * 16384 * 2 + B * T64 * 64 * 4 + B * 4
off = ((ws.data_ptr() + off + 127) & ~127) - ws.data_ptr()
off_cnt = off + B * N * 96 * 2
cnt = ws[off_cnt:off_cnt + B * N * 4].view(torch.int32)
print("survivors: overflow rows", (cnt < 0).sum().item(), "of", cnt.numel(), "mean cnt", cnt[cnt >= 0].float().mean().item(), "max", cnt.max().item())
sys.path.insert(0, "/root/repo/tools")
from kbench import timeit
print("knn_feat on real fm0, B=16: ms", timeit(lambda: ops.knn_feat(fm0, 20)))
os.environ["X"]="1"
