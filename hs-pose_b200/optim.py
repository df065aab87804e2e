"""Optimiser step on the flat parameter / gradient buffers (K9, csrc/optim.cu).

`FlatOptimizer` replaces, for the train-step engine, the reference's
`torch.nn.utils.clip_grad_norm_(params, 5)` + `Ranger.step()` pair
(engine/train.py:105-110; tools/torch_utils/solver/ranger2020.py:135-246, built by
tools/solver_utils.py:49-50) — or Adam, which BASELINE.json configs[2] names — with three launches
over ONE flat fp32 buffer.  The step counter and the learning rate live in device memory, so the
call is captured once inside the step's CUDA graph; `set_lr` (what an LR scheduler calls) is a plain
device write outside the graph.
"""
import ctypes

import torch

from . import _lib, ops

ADAM, RANGER = 0, 1


class FlatOptimizer:
    def __init__(self, flat, kind="ranger", lr=1e-4, betas=None, eps=None, weight_decay=0.0, clip=5.0,
                 alpha=0.5, k=6, n_sma_threshold=5):
        """flat: parallel.FlatGradients after flatten_params() (flat.flat_param / flat.flat / flat.params)."""
        if flat.flat_param is None:
            flat.flatten_params()
        self.kind = {"adam": ADAM, "ranger": RANGER}[kind]
        self.flat = flat
        self.p, self.g = flat.flat_param.data, flat.flat
        if not self.p.is_cuda:
            raise _lib.HSPoseLibraryError("FlatOptimizer: parameters must live on a CUDA device (no CPU path)")
        dev = self.p.device
        # defaults: Ranger as the reference constructs it (ranger2020.py:47-53), Adam as torch.optim.Adam
        self.betas = betas or ((0.95, 0.999) if self.kind == RANGER else (0.9, 0.999))
        self.eps = eps if eps is not None else (1e-5 if self.kind == RANGER else 1e-8)
        self.weight_decay, self.clip, self.alpha, self.k, self.nsma = weight_decay, clip, alpha, k, n_sma_threshold
        self.exp_avg = torch.zeros_like(self.p)
        self.exp_avg_sq = torch.zeros_like(self.p)
        self.slow = self.p.clone() if self.kind == RANGER else None   # ranger2020.py:162-163
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        off, ln = [], []
        for prm, o in zip(flat.params, flat.offsets):
            n = prm.numel()
            if self.kind == RANGER and prm.dim() > 1:          # gradient centralisation per output row
                rows, rl = prm.shape[0], n // prm.shape[0]
                off += [o + r * rl for r in range(rows)]
                ln += [rl] * rows
            else:
                for c in range(0, n, 1024):
                    off.append(o + c)
                    ln.append(-min(1024, n - c))
        self.seg_off = torch.tensor(off, dtype=torch.int32, device=dev)
        self.seg_len = torch.tensor(ln, dtype=torch.int32, device=dev)
        lib = _lib.load()
        self.ws = torch.empty(int(lib.hsp_optim_workspace_bytes()), dtype=torch.uint8, device=dev)

    def set_lr(self, lr):
        self.lr.fill_(float(lr))

    def step(self):
        """Clip (if configured) and update; returns the device scalar ||g|| before clipping."""
        f = ctypes.c_float
        with torch.cuda.device(self.p.device):
            ops._call("hsp_optim_step", self.kind, ops._p(self.p), ops._p(self.g), ops._p(self.exp_avg),
                      ops._p(self.exp_avg_sq), ops._p(self.slow), ctypes.c_long(self.p.numel()), ops._p(self.seg_off),
                      ops._p(self.seg_len), int(self.seg_off.numel()), ops._p(self.step_count), ops._p(self.lr),
                      f(self.betas[0]), f(self.betas[1]), f(self.eps), f(self.weight_decay),
                      f(self.clip if self.clip else 0.0), f(self.alpha), self.k, self.nsma, ops._p(self.grad_norm),
                      ops._p(self.ws), self.ws.numel(), ops._stream())
        return self.grad_norm

    # ---- engine.TrainStep's side-effect-free warm-up and checkpointing
    def state_tensors(self):
        t = {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_count, "lr": self.lr}
        if self.slow is not None:
            t["slow_buffer"] = self.slow
        return t

    def state_dict(self):
        return {k: v.clone() for k, v in self.state_tensors().items()}

    def load_state_dict(self, sd):
        for k, v in self.state_tensors().items():
            v.copy_(sd[k])
