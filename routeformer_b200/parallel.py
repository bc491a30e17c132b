"""Data-parallel training step: one process per GPU, clips sharded over ranks, ONE gradient collective.

Replaces Lightning's DDPStrategy(find_unused_parameters=True, process_group_backend="nccl") of
experiments/full_comparison.py:794 and the AdamW + clip_grad_norm(2.5) of :694-702,829-830:
  * every trainable parameter's gradient is a view into the flat gradient arena, zero-filled once per step, so modules
    skipped by view/gaze dropout simply contribute zeros (no find_unused_parameters bookkeeping);
  * gradients are summed across ranks with NCCL all-reduce over the arena in a few large buckets (NVLink/NVSwitch: the cost
    is launch latency, not link count) -- the only collective on the path;
  * the optimiser is one fused AdamW launch over the arena that also applies the 1/world scaling and the global-norm clip.
BatchNorm statistics stay per-rank, as under the reference's plain DDP.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from .arena import Arena


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Contiguous shard of the global batch for `rank` (global batch 512 -> 256/128/64 clips on 2/4/8 GPUs)."""
    out = {}
    for k, v in batch.items():
        B = v.shape[0]
        if B % world:
            raise ValueError(f"global batch {B} is not divisible by world size {world}")
        per = B // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def _clone_tree(x):
    """Deep copy of a (possibly nested) dict / tuple / list of tensors, keeping StagedBatch's `video_len`."""
    if isinstance(x, torch.Tensor):
        return x.clone()
    if isinstance(x, dict):
        out = type(x)()
        for k, v in x.items():
            out[k] = _clone_tree(v)
        if hasattr(x, "video_len"):
            out.video_len = dict(x.video_len)
        return out
    if isinstance(x, (tuple, list)):
        return type(x)(_clone_tree(v) for v in x)
    return x


def _copy_tree(dst, src) -> None:
    if isinstance(dst, torch.Tensor):
        if src.data_ptr() != dst.data_ptr():
            dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_tree(dst[k], src[k])
    elif isinstance(dst, (tuple, list)):
        for d, s_ in zip(dst, src):
            _copy_tree(d, s_)


def bucket_bounds(n: int, n_buckets: int, align: int = 1024) -> List[Tuple[int, int]]:
    """Splits [0, n) into <= n_buckets aligned ranges, last bucket first (reverse execution order: the Informer decoder's
    gradients, which autograd produces first, sit at the end of the arena)."""
    if n <= 0:
        return []
    size = max(align, (n + n_buckets - 1) // n_buckets)
    size = (size + align - 1) // align * align
    bounds = [(lo, min(n, lo + size)) for lo in range(0, n, size)]
    return bounds[::-1]


def allreduce_flat(flat: torch.Tensor, group=None, n_buckets: int = 4) -> List:
    """Sum-all-reduce of a flat gradient buffer in a few large buckets; returns the async work handles."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    return [dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True) for lo, hi in bucket_bounds(flat.numel(), n_buckets)]


class DataParallelTrainer:
    """`use_cuda_graph=True` captures zero-grad + forward + loss + backward (~950 kernel launches) into ONE CUDA graph after
    two eager warm-up steps and replays it every step: the launch-bound tail of small kernels no longer waits on the host.
    Requirements: static shapes, dropouts 0 (no data-dependent control flow); the CPU random draws of every step are still
    made on the host in reference order and reach the graph through a persistent pinned buffer."""

    def __init__(self, model, loss_fn: Callable, lr: float = 1e-5, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 2.5, group=None, n_buckets: int = 4, use_cuda_graph: bool = False,
                 overlap_wgrad: bool = True, step_fn: Optional[Callable] = None, draw_fn: Optional[Callable] = None,
                 overlap_allreduce: bool = True):
        """step_fn(batch, targets) -> loss (optional) replaces `loss_fn(model(batch), targets)`: e.g. the reference's whole
        `training_step` (forward + eval-mode target pass + both losses, `ParallelTrainerSteps.graph_hooks`).  draw_fn(batch,
        refill_only) -> drop pattern must then make ALL CPU random draws of that step, in order (used to refresh the pinned index
        buffers before a replay).  `graph_tag` (any hashable) is part of the graph cache key: set it when the step function
        changes (e.g. the dense-loss weight switching on after epoch 10)."""
        self.model = model
        self.loss_fn = loss_fn
        self.step_fn, self.draw_fn = step_fn, draw_fn
        self.graph_tag = None
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.lr, self.wd, self.betas, self.eps, self.max_grad_norm = lr, weight_decay, betas, eps, max_grad_norm
        self.n_buckets = n_buckets
        self.arena = Arena.ensure(model)
        n = self.arena.n_trainable
        dev = self.arena.device
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_count = 0
        self.use_cuda_graph = use_cuda_graph
        self.wgrad_stream = torch.cuda.Stream(dev) if (overlap_wgrad and dev.type == "cuda") else None
        # Early all-reduce (the reference's DDP buckets overlap with backward, full_comparison.py:794): the GPS backbone holds
        # 78 % of the gradient bytes and finishes its backward pass first, so its slices of the arena are all-reduced on a side
        # stream -- inside the captured step -- while the visual encoders' backward pass is still running.
        self.comm_stream = torch.cuda.Stream(dev) if (overlap_allreduce and self.world > 1 and dev.type == "cuda") else None
        self.early_ranges, self.late_ranges = self._split_ranges("gps_backbone.") if self.comm_stream is not None else ([], [])
        self._early_work = []
        self._graph = None       # most recently captured graph (None until the first graph step)
        self._graphs = {}        # drop pattern -> (graph, static loss tensor, launches)
        self._pool = None
        self._next_pattern = (False, False, False)
        self.pattern_counts = {}
        self.static_batch = None
        self.static_targets = None
        self._replayed = None
        self.graph_launches = 0

    def _split_ranges(self, prefix: str):
        """(ranges of the gradient arena holding the parameters named `prefix*`, the complementary ranges), merged and sorted."""
        arena = self.arena
        spans = sorted((arena.offsets[id(p)], arena.offsets[id(p)] + p.numel()) for n, p in self.model.named_parameters()
                       if p.requires_grad and n.startswith(prefix))
        early = []
        for lo, hi in spans:
            if early and lo - early[-1][1] < 2 * 64:  # alignment padding between groups (always zero): keep the range whole
                early[-1][1] = hi
            else:
                early.append([lo, hi])
        late, pos = [], 0
        for lo, hi in early:
            if lo > pos:
                late.append((pos, lo))
            pos = hi
        if pos < arena.n_trainable:
            late.append((pos, arena.n_trainable))
        return [tuple(r) for r in early], late

    def _early_allreduce(self) -> None:
        """Backward hook: the GPS backbone's gradients are complete (their kernels are enqueued on the main / wgrad streams)."""
        main = torch.cuda.current_stream()
        self.comm_stream.wait_stream(main)
        if self.wgrad_stream is not None:
            self.comm_stream.wait_stream(self.wgrad_stream)
        with torch.cuda.stream(self.comm_stream):
            for lo, hi in self.early_ranges:
                self._early_work.append(dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # -- forward + backward ----------------------------------------------------------------------
    def _fwd_bwd(self, batch, targets) -> torch.Tensor:
        from . import functional as Fn

        self.arena.zero_grad()
        early = self.comm_stream is not None
        self.model.backbone_grads_ready_hook = self._early_allreduce if early else None
        self._early_work = []
        try:
            if self.step_fn is not None:
                loss = self.step_fn(batch, targets)
            else:
                out = self.model(batch)
                loss = self.loss_fn(out, targets)
            main = torch.cuda.current_stream()
            if self.wgrad_stream is None:
                loss.backward()
            else:
                # weight-gradient GEMMs on a side stream, concurrent with the dgrad chain; joined before anyone reads the gradients
                self.wgrad_stream.wait_stream(main)  # the zero-fill of the gradient arena precedes the first accumulation
                Fn.WgradStream.stream = self.wgrad_stream
                try:
                    loss.backward()
                finally:
                    Fn.WgradStream.stream = None
                    main.wait_stream(self.wgrad_stream)
        finally:
            self.model.backbone_grads_ready_hook = None
        if early:
            if not self._early_work:
                raise RuntimeError("the early all-reduce did not fire: the GPS backbone's input carried no gradient")
            with torch.cuda.stream(self.comm_stream):
                for w in self._early_work:
                    w.wait()
            main.wait_stream(self.comm_stream)
            self._early_work = []
        return loss

    def _ensure_static(self, batch, targets) -> None:
        if self.static_batch is not None:
            return
        # Dedicated static input buffers: the graphs never read a tensor somebody else owns (e.g. a BatchPrefetcher ring slot
        # that the side-stream staging of a later step overwrites while a replay is reading it).  `static_inputs()` hands them
        # out for callers that want to fill them in place and skip the per-step device copy.
        self.static_batch, self.static_targets = _clone_tree(batch), _clone_tree(tuple(targets))

    def _draw(self, refill_only):
        """All CPU random draws of one step (reference order) -> drop pattern of the training forward."""
        if self.draw_fn is not None:
            return self.draw_fn(self.static_batch, refill_only)
        self.model.prepare_draws(self.static_batch, refill_only=refill_only)
        return self.model.last_pattern

    def _capture(self, batch, targets, pattern=None) -> None:
        """Captures zero-grad + forward + loss + backward for one view/gaze drop pattern (routeformer.py:299-310,405-408: dropped
        views skip their whole frame-encoder pass, so every pattern is a different launch sequence).  pattern = (drop_left,
        drop_right, drop_gaze); None = whatever the model draws (dropouts off: always (False, False, False))."""
        m = self.model
        if m.training and m.motion_noise > 0:
            raise ValueError("CUDA-graph capture needs a static step: motion noise must be 0")
        self._ensure_static(batch, targets)
        batch, targets = self.static_batch, self.static_targets
        # The two eager warm-up passes and the capture pass are real training forwards without an optimiser step: without the
        # snapshot below they would leave three momentum updates in the BatchNorm running statistics and advance the CPU RNG
        # stream by three draw plans, so the graph path would diverge from the eager / reference path from step 1.
        rng_state = torch.get_rng_state()
        buffers = {name: b.clone() for name, b in m.named_buffers() if "running_" in name or name.endswith("num_batches_tracked")}
        drop_state = ops.DropoutStream.snapshot(self.arena.device)  # the warm-up forwards advance the dropout step counter
        deferred, m._deferred_tables = m._deferred_tables, []  # this step's own draws, made before the capture was found necessary
        key = (pattern if pattern is not None else (False, False, False), self.graph_tag)
        m.forced_pattern = pattern
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._fwd_bwd(batch, targets)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = ops.launch_count
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(graph, pool=self._pool):
                loss = self._fwd_bwd(batch, targets)
            self._graphs[key] = (graph, loss, ops.launch_count - before)
            self.graph_launches = ops.launch_count - before  # kernels of this library inside the captured step
        finally:
            m.forced_pattern = None
            m._deferred_tables = deferred
        with torch.no_grad():
            for name, b in m.named_buffers():
                if name in buffers:
                    b.copy_(buffers[name])
        torch.set_rng_state(rng_state)
        ops.DropoutStream.restore(self.arena.device, drop_state)
        torch.cuda.synchronize()
        self._graph = graph

    def capture_patterns(self, batch, targets, patterns=None) -> None:
        """Captures the graphs of the given drop patterns up front (default: every pattern the configured dropouts can produce),
        so that no capture falls into a timed region."""
        m = self.model
        if patterns is None:
            views = [(False, False)] + ([(True, False), (False, True)] if (m.training and m.view_dropout > 0 and m.with_scene) else [])
            gazes = [False] + ([True] if (m.training and m.gaze_dropout > 0 and m.with_gaze) else [])
            patterns = [(dl, dr, dg) for dl, dr in views for dg in gazes]
        for pat in patterns:
            if (pat, self.graph_tag) not in self._graphs:
                self._capture(batch, targets, pat)

    def static_inputs(self):
        """(batch, targets) the captured graphs read (None before the first graph step): fill them in place to avoid the copy."""
        return self.static_batch, self.static_targets

    def _replay(self, batch, targets) -> torch.Tensor:
        self._ensure_static(batch, targets)
        if batch is not self.static_batch:
            _copy_tree(self.static_batch, batch)
        if targets is not self.static_targets:
            _copy_tree(self.static_targets, tuple(targets))
        m = self.model
        if not m._deferred_tables:  # draws not made ahead of time by prefetch_draws(): make them now (CPU RNG, reference order)
            self._next_pattern = self._draw("defer")
        pattern = self._next_pattern
        if (pattern, self.graph_tag) not in self._graphs:
            if self.comm_stream is not None and self._graphs:
                # a capture runs two eager warm-up steps whose collectives the other ranks would not match
                raise RuntimeError("with the overlapped all-reduce every drop pattern must be captured up front, on all ranks in the "
                                   "same order: call capture_patterns(batch, targets) before the first step")
            self._capture(self.static_batch, self.static_targets, pattern)
        if self._replayed is not None:
            self._replayed.synchronize()  # the previous replay has consumed the pinned index buffers
        m.commit_deferred_draws()
        graph, loss, launches = self._graphs[(pattern, self.graph_tag)]
        self.graph_launches = launches
        graph.replay()
        if m.training and getattr(m, "uses_dropout", False):
            ops.DropoutStream.note_replay(self.arena.device)
        self._replayed = torch.cuda.Event()
        self._replayed.record()
        self.pattern_counts[pattern] = self.pattern_counts.get(pattern, 0) + 1
        return loss

    def prefetch_draws(self) -> None:
        """Optional: makes the NEXT step's CPU random draws now (same order of the CPU RNG stream), while the current replay is
        still running on the device, so that the next `step` only has to copy them into the pinned buffer and launch."""
        if self._graphs and not self.model._deferred_tables:
            self._next_pattern = self._draw("defer")

    def broadcast_parameters(self, src: int = 0) -> None:
        if self.world > 1:
            dist.broadcast(self.arena.param, src=src, group=self.group)

    def step(self, batch, targets) -> torch.Tensor:
        """fwd + loss + bwd + gradient all-reduce + clip + AdamW.  Returns the (local) loss tensor."""
        arena = Arena.ensure(self.model)
        if arena is not self.arena:
            raise RuntimeError("the model's parameter storage changed after the trainer was built")
        if not self.use_cuda_graph:
            loss = self._fwd_bwd(batch, targets)
        else:
            loss = self._replay(batch, targets)
        if self.comm_stream is not None:  # the backbone's slices were all-reduced inside the step, overlapped with backward
            work = [dist.all_reduce(arena.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for lo, hi in self.late_ranges]
        else:
            work = allreduce_flat(arena.grad, self.group, self.n_buckets)
        for w in work:
            w.wait()
        self.step_count += 1
        self.gnorm_sq.zero_()
        ops.sumsq_accumulate(arena.grad, self.gnorm_sq)
        ops.adamw_step(arena.param[:arena.n_trainable], arena.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1],
                       self.eps, self.wd, self.step_count, 1.0 / self.world, self.gnorm_sq, self.max_grad_norm)
        return loss


class GraphedForward:
    """Inference at speed: the eval-mode forward (~400 launches, most of them a few microseconds long) captured once into a
    CUDA graph and replayed per batch.  The CPU random draws of every forward (ProbSparse key sampling, reference order) are
    still made on the host and reach the graph through the model's persistent pinned index buffer, exactly as in
    DataParallelTrainer.  Static shapes; not for the autoregressive branch (its window loop re-plans draws per window)."""

    def __init__(self, model):
        if model.configs.autoregressive:
            raise ValueError("GraphedForward does not cover the autoregressive branch")
        self.model = model
        self._graph = None
        self.static_batch = None
        self._out = None
        self._replayed = None
        self.graph_launches = 0

    def _capture(self, batch) -> None:
        m = self.model
        static = type(batch)() if isinstance(batch, dict) else {}
        for k, v in batch.items():
            static[k] = v.clone()
        if hasattr(batch, "video_len"):
            static.video_len = dict(batch.video_len)
        self.static_batch = static
        rng_state = torch.get_rng_state()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                m(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        before = ops.launch_count
        with torch.no_grad(), torch.cuda.graph(self._graph):
            self._out = m(static)
        self.graph_launches = ops.launch_count - before
        torch.set_rng_state(rng_state)

    def __call__(self, batch):
        if self.model.training:
            raise RuntimeError("GraphedForward is an inference path: call model.eval() first")
        if self._graph is None:
            self._capture(batch)
        if batch is not self.static_batch:
            for k, v in batch.items():
                if v.data_ptr() != self.static_batch[k].data_ptr():
                    self.static_batch[k].copy_(v, non_blocking=True)
        if self._replayed is not None:
            self._replayed.synchronize()  # the previous replay has consumed the pinned index buffer
        if not self.model.commit_deferred_draws():
            self.model.prepare_draws(self.static_batch, training=False, refill_only=True)
        self._graph.replay()
        self._replayed = torch.cuda.Event()
        self._replayed.record()
        return self._out

    def prefetch_draws(self) -> None:
        if self._graph is not None and not self.model._deferred_tables:
            self.model.prepare_draws(self.static_batch, training=False, refill_only="defer")


class BatchPrefetcher:
    """Stages pinned host batches onto the device on a side stream, `depth` batches ahead of the step that consumes them.

    `submit(host_batch, host_targets)` enqueues the H2D copies of the frames the model consumes (Routeformer.stage_batch) into a
    ring of device buffers; `get()` makes the current stream wait for the oldest staged batch and returns it.  The copy of
    step i+1 then overlaps the compute of step i (PCIe Gen5 moves a 64-clip GEM batch, 529 MB, in ~10 ms)."""

    def __init__(self, model, device, depth: int = 2):
        self.model, self.device, self.depth = model, device, depth
        self.stream = torch.cuda.Stream(device)
        self.slots = [None] * depth      # (StagedBatch, targets)
        self.ready = [None] * depth      # event: staging finished (recorded on the side stream)
        self.free = [None] * depth       # event: consumer finished with the slot (recorded on the consumer stream)
        self.head = self.tail = 0

    def submit(self, host_batch, host_targets) -> None:
        i = self.head % self.depth
        self.head += 1
        with torch.cuda.stream(self.stream):
            if self.free[i] is not None:
                self.stream.wait_event(self.free[i])
            prev = self.slots[i]
            batch = self.model.stage_batch(host_batch, self.device, out=None if prev is None else prev[0])
            if prev is None:
                targets = tuple(t.to(self.device, non_blocking=True) for t in host_targets)
            else:
                targets = prev[1]
                for dst, src in zip(targets, host_targets):
                    dst.copy_(src, non_blocking=True)
            self.slots[i] = (batch, targets)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.ready[i] = ev
        self._last = i

    def get(self):
        i = self.tail % self.depth
        self.tail += 1
        torch.cuda.current_stream().wait_event(self.ready[i])
        return self.slots[i]

    def release(self, slot_batch) -> None:
        """Call after the step that consumed `slot_batch` has been enqueued: its buffers may be overwritten once that work is done."""
        for i, s in enumerate(self.slots):
            if s is not None and s[0] is slot_batch:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                self.free[i] = ev
