"""Data parallelism over the GPUs of one box: one process per GPU, the batch
axis sharded, ONE all-reduce of the gradients per step.

The reference is single-GPU (engine/train.py:23, script.sh:1) — there is
nothing to port.  Objects are independent in every hot-path op, so forward and
backward need no collective; only the 9,709,871 fp32 gradients (38.8 MB) are
summed over NVLink/NVSwitch with NCCL and divided by the world size, before
`clip_grad_norm_` (the clip must see the reduced gradient, engine/train.py:107).

`FlatGradients` re-homes every `param.grad` as a view into one contiguous
buffer, so the collective is a single in-place NCCL call with no pack/unpack
copies.  BatchNorm statistics stay per replica (the reference has no SyncBN) and
every rank must draw the same `Pool_layer` permutation: seed the CPU generator
identically on all ranks (`seed_all`).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def seed_all(seed):
    """Same CPU-generator seed on every rank: Pool_layer's randperm (reference
    gcn3d.py:243) must select the same rows everywhere to equal one big batch."""
    torch.manual_seed(seed)


def shard_batch(batch, rank, world):
    """Contiguous split of every (B, ...) tensor along the batch axis."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v) and v.dim() > 0:
            per = v.shape[0] // world
            out[k] = v[rank * per:(rank + 1) * per]
        else:
            out[k] = v
    return out


class FlatGradients:
    """All gradients of `params` as views of one flat fp32 buffer + one all-reduce."""

    def __init__(self, params, early=None):
        """`early`: optional predicate(param) -> True for parameters whose gradients are complete EARLY in the
        backward (the pose / reconstruction heads: autograd reaches them before the backbone).  They are laid
        out first in the flat buffer, as one contiguous bucket that `all_reduce_early()` can reduce on a side
        stream while the backbone's backward is still running."""
        self.params = [p for p in params if p.requires_grad]
        self.n_early = 0
        if early is not None:
            first = [p for p in self.params if early(p)]
            rest = [p for p in self.params if not early(p)]
            self.params = first + rest
            self.n_early = sum(p.numel() for p in first)
            self.early_params = first
        # every parameter starts on a 32-byte boundary of the flat fp32 buffers (kernels read weights and write
        # gradients with vector accesses) = a 16-byte boundary of the bf16 shadow of the parameters, which TMA
        # descriptors need; the few padding elements stay zero
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 7) // 8 * 8
        n = off
        if self.n_early:
            self.n_early = self.offsets[len(self.early_params)] if len(self.early_params) < len(self.params) else n
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.flat_param = None

    def flatten_params(self):
        """Re-home every parameter's storage as a view of ONE flat buffer (same order as the flat
        gradient) and return it as a single nn.Parameter whose .grad is the flat gradient: the
        optimiser step then is one element-wise kernel over 9.7 M elements instead of one
        multi-tensor launch group per 160 tensors.  Module parameters keep their identity, names and
        values (state_dict is unchanged); only their storage moves."""
        if self.flat_param is None:
            flat = torch.zeros_like(self.flat)
            with torch.no_grad():
                for p, off in zip(self.params, self.offsets):
                    n = p.numel()
                    flat[off:off + n].copy_(p.reshape(-1))
                    p.data = flat[off:off + n].view_as(p)
            self.flat_param = torch.nn.Parameter(flat)
            self.flat_param.grad = self.flat
        return self.flat_param

    def refresh_shadow(self):
        """bf16 copy of ALL parameters in one launch (needs flatten_params()): the tensor-core GEMMs of the
        mixed-precision step read their weights as views of this buffer (ops.weight_shadow) instead of casting
        ~40 weight tensors one by one every step."""
        if self.flat_param is None:
            return None
        if getattr(self, "shadow16", None) is None:
            self.shadow16 = torch.empty_like(self.flat_param.data, dtype=torch.bfloat16)
        self.shadow16.copy_(self.flat_param.data)
        return self.shadow16

    def zero(self):
        """zero_grad() that keeps the views (never set_to_none)."""
        self.flat.zero_()

    def _reduce(self, t):
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(self.world)

    def all_reduce_mean(self, skip_early=False):
        """Average the flat gradient over the ranks (ONE collective, or the late bucket only when the early
        bucket has already been launched by all_reduce_early)."""
        if self.world == 1:
            return
        self._reduce(self.flat[self.n_early:] if (skip_early and self.n_early) else self.flat)

    def all_reduce_early(self):
        """Average the early bucket (the heads' gradients)."""
        if self.world > 1 and self.n_early:
            self._reduce(self.flat[:self.n_early])

    def clip_(self, max_norm):
        """clip_grad_norm_ on the flat buffer (one norm kernel instead of 160)."""
        total = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total
