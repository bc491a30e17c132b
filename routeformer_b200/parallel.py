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
    def __init__(self, model, loss_fn: Callable, lr: float = 1e-5, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 2.5, group=None, n_buckets: int = 4):
        self.model = model
        self.loss_fn = loss_fn
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

    def broadcast_parameters(self, src: int = 0) -> None:
        if self.world > 1:
            dist.broadcast(self.arena.param, src=src, group=self.group)

    def step(self, batch, targets) -> torch.Tensor:
        """fwd + loss + bwd + gradient all-reduce + clip + AdamW.  Returns the (local) loss tensor."""
        arena = Arena.ensure(self.model)
        if arena is not self.arena:
            raise RuntimeError("the model's parameter storage changed after the trainer was built")
        arena.zero_grad()
        out = self.model(batch)
        loss = self.loss_fn(out, targets)
        loss.backward()
        for w in allreduce_flat(arena.grad, self.group, self.n_buckets):
            w.wait()
        self.step_count += 1
        self.gnorm_sq.zero_()
        ops.sumsq_accumulate(arena.grad, self.gnorm_sq)
        ops.adamw_step(arena.param[:arena.n_trainable], arena.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1],
                       self.eps, self.wd, self.step_count, 1.0 / self.world, self.gnorm_sq, self.max_grad_norm)
        return loss
