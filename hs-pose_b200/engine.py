"""Train-step engine: the whole step (forward, losses, backward, gradient all-reduce,
clip, optimiser) captured once into a CUDA graph and replayed.

The reference's loop (engine/train.py:74-113) launches ~1.5k small kernels per step from
Python; at batch 128 on a B200 that CPU launch path is as long as the GPU work itself.  The
step is shape-static (B, N, k fixed), so it is captured with `torch.cuda.graph` and each
iteration becomes: copy the 12 input tensors into static buffers, refresh the two pooling
permutations (still drawn by `torch.randperm` on the CPU generator, reference gcn3d.py:243),
one graph launch.  Hand-written kernels enter the graph like any other launch (they run on
`torch.cuda.current_stream()`); the NCCL all-reduce is captured too.
"""
import torch

from . import gcn3d, ops, parallel


class TrainStep:
    def __init__(self, model, lr=1e-4, clip=5.0, amp=True, graph=True, optimizer=None,
                 tf32=True):
        self.model, self.clip, self.amp, self.use_graph = model, clip, amp, graph
        self.flat = parallel.FlatGradients(model.posenet.parameters())
        # default optimiser: Adam over ONE flat parameter (see FlatGradients.flatten_params)
        self.opt = optimizer or torch.optim.Adam([self.flat.flatten_params()], lr=lr, fused=True,
                                                 capturable=graph)
        self.graph = None
        self.static_batch = None
        self.static_loss = None
        self.pool_rows = []          # static device buffers, one per Pool_layer call
        self._pool_call, self._draw_inline = 0, True
        self.launches_per_step = None
        if tf32 and amp:
            torch.backends.cuda.matmul.allow_tf32 = True   # fp32 leftovers (K=3 STE, per-object GEMVs)

    # ---- pooling permutations through static buffers
    def _provider(self, vertice_num, pool_num, device):
        i = self._pool_call
        self._pool_call += 1
        if i == len(self.pool_rows):
            self.pool_rows.append((vertice_num, torch.empty(pool_num, dtype=torch.int32, device=device),
                                   torch.empty(pool_num, dtype=torch.int32).pin_memory()))
        vertice_num, dev_buf, host_buf = self.pool_rows[i]
        if self._draw_inline:      # eager: draw at the point of use, exactly like the reference
            host_buf.copy_(torch.randperm(vertice_num)[:pool_num])
            dev_buf.copy_(host_buf, non_blocking=True)
        return dev_buf

    def _refresh_pool_rows(self):
        """Graph mode: the draws of one forward, in forward order, before the replay."""
        for vertice_num, dev_buf, host_buf in self.pool_rows:
            host_buf.copy_(torch.randperm(vertice_num)[:host_buf.numel()])
            dev_buf.copy_(host_buf, non_blocking=True)

    # ---- the step itself
    def _body(self, batch, draw_inline):
        self._pool_call, self._draw_inline = 0, draw_inline
        prev = gcn3d.set_pool_rows_provider(self._provider)
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                _, losses = self.model(**batch, do_loss=True)
            total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
            self.flat.zero()
            total.backward()
        finally:
            gcn3d.set_pool_rows_provider(prev)
        self.flat.all_reduce_mean()
        if self.clip:
            self.flat.clip_(self.clip)
        self.opt.step()
        return total.detach()

    def _capture(self, batch):
        dev = next(self.model.parameters()).device
        self.static_batch = {k: (torch.empty_like(v, device=dev) if torch.is_tensor(v) else v)
                             for k, v in batch.items()}
        self._load(batch)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body(self.static_batch, draw_inline=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._refresh_pool_rows()
        l0 = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._body(self.static_batch, draw_inline=False)
        self.launches_per_step = ops.launch_count() - l0

    def _load(self, batch):
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static_batch[k].copy_(v, non_blocking=True)

    def __call__(self, batch):
        """One train step.  `batch`: dict of HSPose.forward kwargs (host pinned or device
        tensors).  Returns the (device, detached) total loss of this step."""
        if not self.use_graph:
            dev = next(self.model.parameters()).device
            return self._body({k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v)
                               for k, v in batch.items()}, draw_inline=True)
        if self.graph is None:
            self._capture(batch)
        self._load(batch)
        self._refresh_pool_rows()
        self.graph.replay()
        return self.static_loss
