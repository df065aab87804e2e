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
import contextlib
import copy
import os

import torch

from . import gcn3d, geom, ops, optim, parallel

_WEIGHT_SHADOW = os.environ.get("HSP_WEIGHT_SHADOW", "1") != "0"   # A/B switch for the bf16 parameter shadow
_RING = 4   # pinned staging slots per Pool_layer permutation (bounds how far the CPU may run ahead)


class _PoolRows:
    """Device buffer of one Pool_layer's sampled rows, fed from a small ring of pinned host buffers.
    A slot is rewritten only after the event recorded behind its last H2D copy has completed, so a
    copy still queued behind earlier graph replays never reads a permutation drawn for a later step."""

    def __init__(self, vertice_num, pool_num, device):
        self.vertice_num, self.pool_num = vertice_num, pool_num
        self.dev = torch.empty(pool_num, dtype=torch.int32, device=device)
        self.host = [torch.empty(pool_num, dtype=torch.int32).pin_memory() for _ in range(_RING)]
        self.done = [None] * _RING
        self.i = 0

    def draw(self):
        """torch.randperm on the CPU generator, exactly the reference's call (gcn3d.py:243)."""
        k = self.i % _RING
        self.i += 1
        if self.done[k] is not None:
            self.done[k].synchronize()
        self.host[k].copy_(torch.randperm(self.vertice_num)[:self.pool_num])
        self.dev.copy_(self.host[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.done[k] = ev


class TrainStep:
    def __init__(self, model, lr=1e-4, clip=5.0, amp=True, graph=True, optimizer=None,
                 tf32=False):
        self.model, self.clip, self.amp, self.use_graph = model, clip, amp, graph
        # heads first: their gradients are final long before the backbone's backward ends, so their share of the
        # all-reduce (~70 % of the 38.8 MB) overlaps it on a side stream (only when there is more than one rank)
        backbone = {id(p) for n, p in model.posenet.named_parameters()
                    if n.startswith("face_recon.conv_") or n.startswith("face_recon.bn") or n.startswith("face_recon.pool")}
        self.flat = parallel.FlatGradients(model.posenet.parameters(), early=lambda p: id(p) not in backbone)
        self._overlap = self.flat.world > 1 and self.flat.n_early > 0
        self._side, self._pending, self._early_done = None, 0, None
        if self._overlap:
            for p in self.flat.early_params:
                p.register_post_accumulate_grad_hook(self._on_early_grad)
        # optimiser: "adam" (BASELINE.json configs[2]) or "ranger" (the reference's, tools/solver_utils.py:49-50)
        # = the K9 kernels over ONE flat parameter buffer with the clip folded in; a torch.optim.Optimizer
        # instance is also accepted (then `clip` runs as a separate pass)
        if optimizer is None or isinstance(optimizer, str):
            self.opt = optim.FlatOptimizer(self.flat, kind=optimizer or "adam", lr=lr, clip=clip)
        else:
            self.opt = optimizer
        self.graph = None
        self.static_batch = None
        self.static_loss = None
        self.pool_rows = []          # static device buffers, one per Pool_layer call
        self._pool_call, self._draw_inline = 0, True
        self.launches_per_step = None
        self.tf32 = bool(tf32 and amp)   # opt-in, and scoped to the step body (never left set process-wide)
        self._inputs_copied = None       # event behind the last H2D copy of a batch

    # ---- pooling permutations through static buffers
    def _provider(self, vertice_num, pool_num, device):
        i = self._pool_call
        self._pool_call += 1
        if i == len(self.pool_rows):
            self.pool_rows.append(_PoolRows(vertice_num, pool_num, device))
        rows = self.pool_rows[i]
        if self._draw_inline:      # eager: draw at the point of use, exactly like the reference
            rows.draw()
        return rows.dev

    def _refresh_pool_rows(self):
        """Graph mode: the draws of one forward, in forward order, before the replay."""
        for rows in self.pool_rows:
            rows.draw()

    @contextlib.contextmanager
    def _tf32_scope(self):
        if not self.tf32:
            yield
            return
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            yield
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    # ---- early gradient bucket: all-reduce on a side stream as soon as the last head gradient has landed
    def _on_early_grad(self, p):
        # runs on an autograd worker thread, whose "current stream" is not the step's: work on / fork from the
        # stream the step itself runs on (recorded by _body; the capture stream in graph mode)
        self._pending -= 1
        if self._pending == 0:
            n = len(self.flat.early_params)
            with torch.cuda.stream(self._main):
                self._gather(0, n)                # the heads' gradients -> their slice of the flat buffer
            if self._side is None:
                self._side = torch.cuda.Stream(device=p.device)
            self._side.wait_stream(self._main)
            with torch.cuda.stream(self._side):
                self.flat.all_reduce_early()
            self._early_done = True

    def _gather(self, lo, hi):
        """Autograd handed over one gradient tensor per parameter (`.grad` was None): copy those of parameters
        lo..hi into the flat buffer with ONE multi-tensor copy (no per-parameter `grad += g` launch, ~160 of
        them) and point `.grad` back at the flat views."""
        ps, vs = self.flat.params[lo:hi], self._views[lo:hi]
        got = [(v, p.grad) for v, p in zip(vs, ps) if p.grad is not None]
        if got:
            torch._foreach_copy_([v for v, _ in got], [g for _, g in got])
        for v, p in zip(vs, ps):
            p.grad = v

    # ---- the step itself
    def _body(self, batch, draw_inline):
        self._pool_call, self._draw_inline = 0, draw_inline
        self._pending, self._early_done = (len(self.flat.early_params) if self._overlap else -1), False
        self._main = torch.cuda.current_stream()
        prev = gcn3d.set_pool_rows_provider(self._provider)
        try:
            # mixed precision: ONE bf16 cast of the flat parameter buffer per step; the GEMMs read views of it
            shadow = self.flat.refresh_shadow() if (self.amp and _WEIGHT_SHADOW) else None
            with self._tf32_scope(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp), \
                    ops.weight_shadow(self.flat.flat_param, shadow):
                _, losses = self.model(**batch, do_loss=True)
            total = getattr(losses, "total", None)      # HSPose's fused loss path sums its terms in one reduction
            if total is None:
                total = sum(v.reshape(()).float() for grp in losses.values() for v in grp.values())
            self.flat.zero()
            self._views = [p.grad for p in self.flat.params]
            for p in self.flat.params:
                p.grad = None
            total.backward()
            n_early = len(self.flat.early_params) if (self._overlap and self._early_done) else 0
            self._gather(n_early, len(self.flat.params))
        finally:
            gcn3d.set_pool_rows_provider(prev)
        if self._overlap and self._early_done:
            self.flat.all_reduce_mean(skip_early=True)            # backbone bucket on the main stream
            torch.cuda.current_stream().wait_stream(self._side)   # join the heads' all-reduce
        else:
            self.flat.all_reduce_mean()
        if isinstance(self.opt, optim.FlatOptimizer):
            self.opt.step()                       # clip + update, three launches
        else:
            if self.clip:
                self.flat.clip_(self.clip)
            self.opt.step()
        return total.detach()

    def _capture(self, batch):
        dev = next(self.model.parameters()).device
        self.static_batch = {k: (torch.empty_like(v, device=dev) if torch.is_tensor(v) else v)
                             for k, v in batch.items()}
        self._load(batch)
        # Warm-up (allocator, lazy optimiser state, cuBLAS/NCCL handles) WITHOUT side effects: parameters,
        # optimiser state, BatchNorm buffers and both RNG streams are restored afterwards, so the first
        # captured step is the first update the model sees — graph training follows the same trajectory
        # as the eager loop (and as the reference's, engine/train.py:74-110).
        snap = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body(self.static_batch, draw_inline=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        l0 = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._body(self.static_batch, draw_inline=False)
        self.launches_per_step = ops.launch_count() - l0

    def _snapshot(self):
        dev = next(self.model.parameters()).device
        return {"params": [p.detach().clone() for p in self.model.parameters()],
                "buffers": [b.detach().clone() for b in self.model.buffers()],
                "opt": self.opt.state_dict() if isinstance(self.opt, optim.FlatOptimizer)
                else copy.deepcopy(self.opt.state_dict()),
                "cpu_rng": torch.get_rng_state(), "cuda_rng": torch.cuda.get_rng_state(dev)}

    def _restore(self, snap):
        dev = next(self.model.parameters()).device
        with torch.no_grad():
            for p, q in zip(self.model.parameters(), snap["params"]):
                p.copy_(q)
            for b, q in zip(self.model.buffers(), snap["buffers"]):
                b.copy_(q)
            if isinstance(self.opt, optim.FlatOptimizer):
                self.opt.load_state_dict(snap["opt"])      # copies in place
                self.flat.zero()
                torch.set_rng_state(snap["cpu_rng"])
                torch.cuda.set_rng_state(snap["cuda_rng"], dev)
                return
            # optimiser state IN PLACE (the captured graph holds these tensors' addresses): a state that
            # did not exist before the warm-up (first step) goes back to its initial value, zero
            old = snap["opt"]["state"]
            for idx, st in self.opt.state_dict()["state"].items():
                live = self.opt.state[self._opt_param(idx)]
                for name, val in live.items():
                    if torch.is_tensor(val):
                        if idx in old and name in old[idx]:
                            val.copy_(old[idx][name])
                        else:
                            val.zero_()
            self.flat.zero()
        torch.set_rng_state(snap["cpu_rng"])
        torch.cuda.set_rng_state(snap["cuda_rng"], dev)

    def _opt_param(self, idx):
        flat = [p for g in self.opt.param_groups for p in g["params"]]
        return flat[idx]

    def _load(self, batch):
        """H2D (or D2D) copies of the step's inputs into the static buffers.  The copies are
        asynchronous: a caller that REWRITES its pinned host tensors between steps must call
        `wait_inputs()` first (or hand in fresh tensors); reusing them unchanged needs nothing."""
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static_batch[k].copy_(v, non_blocking=True)
        self._inputs_copied = torch.cuda.Event()
        self._inputs_copied.record()

    def wait_inputs(self):
        """Block until the last batch handed to __call__ has been copied off the host."""
        if self._inputs_copied is not None:
            self._inputs_copied.synchronize()

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


class EvalRunner:
    """The per-image call of the reference's evaluation loop (evaluation/evaluate.py:91-108): forward of all
    detections of one image (batch = 1..~8 objects) + `generate_RT` + `pred_s = Pred_s + mean_shape`, replayed
    from ONE CUDA graph per batch-size bucket.  The published 38 FPS is exactly this region, latency-bound at
    these batch sizes; a graph replay removes the ~300 kernel-launch gaps.  A batch is padded up to its bucket
    by repeating its first object (objects are independent in eval mode: BatchNorm uses running statistics).
    Pool_layer's sample is still drawn with `torch.randperm` on the CPU generator before each replay
    (reference gcn3d.py:243 draws it even in eval mode)."""

    def __init__(self, model, buckets=(1, 2, 4, 8, 16, 32), graph=True):
        self.model, self.buckets, self.use_graph = model, tuple(sorted(buckets)), graph
        self.graphs = {}          # bucket -> (graph, static inputs, static outputs, pool rows)
        self._rows, self._call_i, self._inline = None, 0, True

    def _provider(self, vertice_num, pool_num, device):
        i = self._call_i
        self._call_i += 1
        if i == len(self._rows):
            self._rows.append(_PoolRows(vertice_num, pool_num, device))
        if self._inline:
            self._rows[i].draw()
        return self._rows[i].dev

    def _forward(self, inp):
        out = self.model(PC=inp["PC"], obj_id=inp["obj_id"], mean_shape=inp["mean_shape"], sym=inp["sym"])
        res = {k: out[k] for k in ("p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s")}
        res["pred_RT"] = geom.generate_RT([res["p_green_R"], res["p_red_R"]], [res["f_green_R"], res["f_red_R"]],
                                          res["Pred_T"], "vec", inp["sym"])
        res["pred_s"] = res["Pred_s"] + inp["mean_shape"]
        return res

    def _run(self, inp, rows, inline):
        self._rows, self._call_i, self._inline = rows, 0, inline
        prev = gcn3d.set_pool_rows_provider(self._provider)
        try:
            with torch.no_grad():
                return self._forward(inp)
        finally:
            gcn3d.set_pool_rows_provider(prev)

    @torch.no_grad()
    def __call__(self, PC, obj_id, mean_shape, sym):
        n = PC.shape[0]
        dev = next(self.model.parameters()).device
        if n == 0:
            return {"pred_RT": torch.zeros(0, 4, 4, device=dev), "pred_s": torch.zeros(0, 3, device=dev)}
        nb = next((b for b in self.buckets if b >= n), None)
        given = {"PC": PC, "obj_id": obj_id, "mean_shape": mean_shape, "sym": sym}
        if not self.use_graph or nb is None:      # larger than the largest bucket: plain eager forward
            rows = []
            return self._run({k: v.to(dev) for k, v in given.items()}, rows, True)
        if nb not in self.graphs:
            inp = {k: torch.zeros((nb,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev) for k, v in given.items()}
            rows = []
            self._fill(inp, given, n, nb)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._run(inp, rows, True)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._run(inp, rows, False)
            self.graphs[nb] = (g, inp, out, rows)
        g, inp, out, rows = self.graphs[nb]
        self._fill(inp, given, n, nb)
        for r in rows:
            r.draw()
        g.replay()
        return {k: v[:n].clone() for k, v in out.items()}

    @staticmethod
    def _fill(inp, given, n, nb):
        for k, v in given.items():
            inp[k][:n].copy_(v, non_blocking=True)
            if nb > n:
                inp[k][n:].copy_(inp[k][:1].expand(nb - n, *inp[k].shape[1:]))
