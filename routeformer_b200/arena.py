"""Flat parameter / gradient arena.

All parameters of a module tree are re-homed as views into ONE fp32 buffer (trainable parameters first), and their
`.grad`s as views into ONE gradient buffer of the same layout.  This is what makes the rest of the design possible:
  * q/k/v projection weights (and biases) of every attention layer sit back to back, so the fused QKV GEMM and its
    wgrad address them as a single [3D, D] matrix while `state_dict()` keeps the reference's per-tensor layout;
  * wgrad GEMMs accumulate straight into the gradient arena (no autograd AccumulateGrad kernels);
  * data parallelism is one NCCL all-reduce over `grad[:n_trainable]`, the optimiser one fused AdamW launch.
`load_state_dict` copies in place, so it preserves the arena; `module.to(device)` re-creates storages, after which
`Arena.ensure(module)` rebuilds it (cheap check per call).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

ALIGN = 64  # elements (256 B): every group starts TMA-aligned


def _groups(module: nn.Module) -> List[List[nn.Parameter]]:
    """Parameters in arena order: per attention layer [Wq,Wk,Wv], [bq,bk,bv]; everything else on its own."""
    from .layers import AttentionLayer

    seen, groups = set(), []
    for m in module.modules():
        if isinstance(m, AttentionLayer):
            ws = [m.query_projection.weight, m.key_projection.weight, m.value_projection.weight]
            bs = [m.query_projection.bias, m.key_projection.bias, m.value_projection.bias]
            for g in (ws, bs):
                if all(id(p) not in seen for p in g) and len({p.requires_grad for p in g}) == 1:
                    groups.append(g)
                    seen.update(id(p) for p in g)
    for p in module.parameters():
        if id(p) not in seen:
            groups.append([p])
            seen.add(id(p))
    groups.sort(key=lambda g: not g[0].requires_grad)  # stable: trainable groups first
    return groups


class Arena:
    def __init__(self, module: nn.Module):
        groups = _groups(module)
        params = [p for g in groups for p in g]
        if not params:
            raise ValueError("module has no parameters")
        dev = params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in params):
            raise ValueError("Arena needs all parameters in fp32 on one device")
        offsets, off, n_trainable = {}, 0, 0
        for g in groups:
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            for p in g:
                offsets[id(p)] = off
                off += p.numel()
            if g[0].requires_grad:
                n_trainable = off
        self.n_trainable = (n_trainable + ALIGN - 1) // ALIGN * ALIGN
        self.size = (off + ALIGN - 1) // ALIGN * ALIGN
        self.device = dev
        self.param = torch.zeros(self.size, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(self.n_trainable, device=dev, dtype=torch.float32)
        self.params = params
        self.offsets = offsets
        with torch.no_grad():
            for p in params:
                o, n = offsets[id(p)], p.numel()
                self.param[o:o + n].copy_(p.data.reshape(-1))
                old_grad = p.grad
                p.data = self.param[o:o + n].view(p.shape)
                if p.requires_grad:
                    g = self.grad[o:o + n].view(p.shape)
                    if old_grad is not None:
                        g.copy_(old_grad)
                    p.grad = g
        self._ptrs = [p.data_ptr() for p in params]

    # -- maintenance ---------------------------------------------------------------------------
    def valid(self) -> bool:
        return all(p.data_ptr() == q for p, q in zip(self.params, self._ptrs))

    def attach_grads(self) -> None:
        """Re-attaches gradient views dropped by `zero_grad(set_to_none=True)` (they come back zeroed)."""
        missing = [p for p in self.params if p.requires_grad and (p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * self.offsets[id(p)])]
        if not missing:
            return
        if len(missing) == sum(p.requires_grad for p in self.params):
            self.grad.zero_()
        for p in missing:
            o, n = self.offsets[id(p)], p.numel()
            g = self.grad[o:o + n].view(p.shape)
            if len(missing) != sum(q.requires_grad for q in self.params):
                g.zero_()
            p.grad = g

    def zero_grad(self) -> None:
        self.grad.zero_()

    @staticmethod
    def ensure(module: nn.Module) -> "Arena":
        arena: Optional[Arena] = module.__dict__.get("_rf_arena")
        if arena is None or not arena.valid() or len(arena.params) != sum(1 for _ in module.parameters()):
            arena = Arena(module)
            module.__dict__["_rf_arena"] = arena
        arena.attach_grads()
        return arena
