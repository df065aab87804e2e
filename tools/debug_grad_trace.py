"""Which backward is not reproducible?  Records the gradient arriving at the output of every ops.* call of one
train forward/backward (tensor hooks), twice, and lists the calls in forward order with the run-to-run difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hspose_b200.ops as ops
from hspose_b200 import parallel
from hspose_b200.HSPose import HSPose
from hspose_b200.flags import get_flags
from hspose_b200.synth import fill_params, synth_batch

NAMES = ["surface_conv", "graph_conv", "hs_conv_mixed", "gather_max", "orl_global", "residual_sum", "linear_tc",
         "linear_bn_relu", "multi_linear_bn_relu", "bn_relu", "colmax", "normalize_dirs", "split_halves",
         "concat_upsample", "upsample_rows", "gather_rows", "chamfer", "fused_losses"]
store, counter = {}, {}


def wrap(name, f):
    def g(*a, **k):
        out = f(*a, **k)
        i = counter.get(name, 0)
        counter[name] = i + 1
        outs = out if isinstance(out, (tuple, list)) else (out,)
        for j, o in enumerate(outs):
            if torch.is_tensor(o) and o.requires_grad:
                key = (len(store_order), name, i, j, tuple(o.shape), str(o.dtype)[6:])
                store_order.append(key)
                o.register_hook(lambda gr, key=key: store.__setitem__(key, gr.detach().clone()))
        return out
    return g


for n in NAMES:
    if hasattr(ops, n):
        setattr(ops, n, wrap(n, getattr(ops, n)))
dev = torch.device("cuda")
F = get_flags()
for n in ("aug_pc_pro", "aug_rt_pro", "aug_bb_pro", "aug_bc_pro"):
    setattr(F, n, 0.0)
model = fill_params(HSPose("PoseNet_only", chamfer_w=1.0)).to(dev).train()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
batch = {k: v.to(dev) for k, v in synth_batch(4, 1028, seed=9, train=True).items()}
runs = []
for r in range(2):
    parallel.seed_all(4321)
    store, counter, store_order = {}, {}, []
    for p in model.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, losses = model(**batch, do_loss=True)
    total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
    total.backward()
    torch.cuda.synchronize()
    runs.append((dict(store), list(store_order)))
a, order = runs[0]
b = {k[1:]: v for k, v in runs[1][0].items()}
print("grad arriving at the OUTPUT of each call (forward order); rel L2 difference run 0 vs run 1")
for key in order:
    if key not in a or key[1:] not in b:
        continue
    x, y = a[key].float(), b[key[1:]].float()
    d = (x - y).norm().item() / (x.norm().item() + 1e-30)
    print(f"{key[0]:4d} {key[1]:22s}#{key[2]:<2d} out{key[3]} {str(key[4]):22s} {key[5]:9s} {d:.3e}")
