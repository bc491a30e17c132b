"""torch.autograd.Function wrappers: each one is a fused block of the Routeformer path whose forward AND
backward are sequences of calls into the CUDA library (ops.py).  PyTorch only allocates buffers and orders
the calls; every weight gradient is accumulated by the wgrad GEMM directly into `param.grad` (which is a
view of the flat gradient arena when the model is managed by `routeformer_b200.arena.Arena`), and the
Function returns None for it, so autograd never launches its own accumulation kernels for parameters.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function

from . import ops


def grad_buffer(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The tensor weight gradients are accumulated into (created zero-filled on first use)."""
    if p is None or not p.requires_grad:
        return None
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class WgradStream:
    """Optional side stream for the weight-gradient GEMMs.  dW = dY^T X only has to be ready when the optimiser runs, while the
    backward chain waits for dX = dY W: with a side stream the two run concurrently, which matters for the small layers (Informer,
    gaze / video encoders) whose GEMMs fill a fraction of the SMs.  Set for the duration of a step by DataParallelTrainer, which
    also joins the stream before touching the gradients; None = everything on the caller's stream."""
    stream: Optional[torch.cuda.Stream] = None


def wgrad(dy_t: torch.Tensor, x: torch.Tensor, g: torch.Tensor) -> None:
    """g[N,K] += dy_t[M,N]^T x[M,K] (split over the long reduction, accumulated straight into the gradient arena)."""
    side = WgradStream.stream
    if side is None:
        ops.gemm(dy_t, x, g, a_mn=True, b_mn=True, accumulate=True)
        return
    side.wait_stream(torch.cuda.current_stream())  # dy_t was produced on the main stream
    dy_t.record_stream(side)
    x.record_stream(side)
    with torch.cuda.stream(side):
        ops.gemm(dy_t, x, g, a_mn=True, b_mn=True, accumulate=True)


def _adjacent(*ts: torch.Tensor) -> bool:
    """True if the (contiguous) tensors sit back to back in ONE storage (the arena), i.e. can be addressed as one matrix."""
    base = ts[0].untyped_storage().data_ptr()
    for a, b in zip(ts[:-1], ts[1:]):
        if not (a.is_contiguous() and b.is_contiguous()) or b.untyped_storage().data_ptr() != base:
            return False
        if a.data_ptr() + a.numel() * 4 != b.data_ptr():
            return False
    return True


def _fused(*ts: torch.Tensor) -> Optional[torch.Tensor]:
    """One [sum rows, cols] view over adjacent row-major matrices / vectors (None if they are not adjacent)."""
    if not _adjacent(*ts):
        return None
    first = ts[0].detach()
    rows = sum(t.shape[0] for t in ts)
    shape = (rows,) + tuple(first.shape[1:])
    return first.as_strided(shape, first.stride(), first.storage_offset())


def _fused_grads(*ps: torch.Tensor) -> Optional[torch.Tensor]:
    gs = [grad_buffer(p) for p in ps]
    if any(g is None for g in gs):
        return None
    return _fused(*gs)


# ------------------------------------------------------------------------------------------------
class AttentionBlock(Function):
    """y = x + OutProj(Attention(Wq x, Wk s, Wv s)),  s = x (self) or `cross`.

    replaces AttentionLayer.forward + the residual add of the calling layer
    (cross_modal_transformer.py:185-198,293-294,224-227; SelfAttentionFamily.py:182-194).
    x [B*Lq, D], cross [B*Lk, D] (row-major 2-D).  `idx` int32 [groups, Lq, U] on the device.
    meta: dict(B, H, Lq, Lk, mode, layout, U, u, idx_group, record)
    """

    @staticmethod
    def forward(ctx, x, cross, wq, bq, wk, bk, wv, bv, wo, bo, idx, meta):
        B, H, Lq, Lk = meta["B"], meta["H"], meta["Lq"], meta["Lk"]
        D = wq.shape[0]
        dh = D // H
        M = x.shape[0]
        self_attn = cross is None
        dev = x.device
        if self_attn:
            qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.float32)
            w_all, b_all = _fused(wq, wk, wv), _fused(bq, bk, bv)
            if w_all is not None and b_all is not None:
                ops.gemm(x, w_all, qkv, bias=b_all)
            else:
                for i, (w, b) in enumerate(((wq, bq), (wk, bk), (wv, bv))):
                    ops.gemm(x, w, qkv[:, i * D:(i + 1) * D], bias=b)
            q = (qkv, Lq * 3 * D, 3 * D)
            k = (qkv[:, D:], Lk * 3 * D, 3 * D)
            v = (qkv[:, 2 * D:], Lk * 3 * D, 3 * D)
            saved_proj = (qkv,)
        else:
            Mk = cross.shape[0]
            qb = torch.empty(M, D, device=dev, dtype=torch.float32)
            kvb = torch.empty(Mk, 2 * D, device=dev, dtype=torch.float32)
            ops.gemm(x, wq, qb, bias=bq)
            w_kv, b_kv = _fused(wk, wv), _fused(bk, bv)
            if w_kv is not None and b_kv is not None:
                ops.gemm(cross, w_kv, kvb, bias=b_kv)
            else:
                ops.gemm(cross, wk, kvb[:, :D], bias=bk)
                ops.gemm(cross, wv, kvb[:, D:], bias=bv)
            q = (qb, Lq * D, D)
            k = (kvb, Lk * 2 * D, 2 * D)
            v = (kvb[:, D:], Lk * 2 * D, 2 * D)
            saved_proj = (qb, kvb)
        context = torch.empty(M, D, device=dev, dtype=torch.float32)
        u = meta["u"]
        top = torch.empty(B, H, max(u, 1), device=dev, dtype=torch.int32) if meta["mode"] != ops.ATTN_FULL else None
        measure = torch.empty(B, H, Lq, device=dev, dtype=torch.float32) if meta.get("record") is not None and top is not None else None
        drop = meta.get("drop")  # training-mode feature dropout: {"p", "out": (seed, off), "prob": (seed, off) | None}
        ops.attention_fwd(q, k, v, B, H, Lq, Lk, dh, meta["mode"], meta["layout"], idx, meta["idx_group"], meta["U"], u, context, top,
                          measure=measure, forced_top=meta.get("forced_top"),
                          dropout=(drop["p"], *drop["prob"]) if drop and drop.get("prob") else None,
                          tail_only=bool(meta.get("tail")) and not drop)
        if meta.get("record") is not None and top is not None:
            meta["record"].append({"where": meta.get("name", ""), "top": top, "measure": measure})
        tail = bool(meta.get("tail")) and not drop
        if tail:
            # only the LAST token of every sequence is consumed downstream (PerceiveEncoder(out_len=1), cross_modal_transformer.py:433):
            # project / add the residual for those B rows only; everything after this block then runs on B instead of B*L rows
            ctx_tail, x_tail = context.view(B, Lq, D)[:, Lq - 1, :], x.view(B, Lq, D)[:, Lq - 1, :]
            y = torch.empty(B, D, device=dev, dtype=torch.float32)
            ops.gemm(ctx_tail, wo, y, bias=bo, residual=x_tail)
        else:
            y = torch.empty(M, D, device=dev, dtype=torch.float32)
            if drop:  # x + dropout(out_projection(context))   (cross_modal_transformer.py:224,227,295)
                ops.gemm(context, wo, y, bias=bo)
                ops.dropout(y, y, drop["p"], *drop["out"], residual=x)
            else:
                ops.gemm(context, wo, y, bias=bo, residual=x)
        ctx.tail = tail
        ctx.meta = meta
        ctx.params = (wq, bq, wk, bk, wv, bv, wo, bo)
        ctx.self_attn = self_attn
        ctx.save_for_backward(x, cross, context, top, *saved_proj)
        return y

    @staticmethod
    def backward(ctx, dy):
        meta = ctx.meta
        B, H, Lq, Lk = meta["B"], meta["H"], meta["Lq"], meta["Lk"]
        wq, bq, wk, bk, wv, bv, wo, bo = ctx.params
        x, cross, context, top, *proj = ctx.saved_tensors
        D = wq.shape[0]
        dh = D // H
        M = x.shape[0]
        dev = x.device
        dy = dy.contiguous()
        drop = meta.get("drop")
        dproj = dy  # gradient of the out-projection output; dy itself stays the gradient of the residual branch
        if drop:
            dproj = ops.dropout(dy, torch.empty_like(dy), drop["p"], *drop["out"])
        attn_drop = (drop["p"], *drop["prob"]) if drop and drop.get("prob") else None
        g = grad_buffer(wo)
        if ctx.tail:  # dy is [B, D]: the gradient of the last-token rows; every other context row has zero gradient
            if g is not None:
                wgrad(dproj, context.view(B, Lq, D)[:, Lq - 1, :], g)
            dcontext = torch.zeros(M, D, device=dev, dtype=torch.float32)
            ops.gemm(dproj, wo, dcontext.view(B, Lq, D)[:, Lq - 1, :], b_mn=True, colsum_a=grad_buffer(bo))
        else:
            if g is not None:
                wgrad(dproj, context, g)
            dcontext = torch.empty(M, D, device=dev, dtype=torch.float32)
            ops.gemm(dproj, wo, dcontext, b_mn=True, colsum_a=grad_buffer(bo))  # bias gradient from the operand tiles of the dgrad
        need_dx = ctx.needs_input_grad[0]
        dx = dcross = None
        if ctx.self_attn:
            (qkv,) = proj
            dqkv = torch.empty_like(qkv)
            q = (qkv, Lq * 3 * D, 3 * D)
            k = (qkv[:, D:], Lk * 3 * D, 3 * D)
            v = (qkv[:, 2 * D:], Lk * 3 * D, 3 * D)
            ops.attention_bwd(q, k, v, B, H, Lq, Lk, dh, meta["mode"], meta["layout"], meta["U"], meta["u"], top, dcontext,
                              dqkv, dqkv[:, D:], dqkv[:, 2 * D:], dropout=attn_drop, tail_only=ctx.tail)
            w_all = _fused(wq, wk, wv)
            gw_all, gb_all = _fused_grads(wq, wk, wv), _fused_grads(bq, bk, bv)
            if gw_all is not None:
                wgrad(dqkv, x, gw_all)
            else:
                for i, w in enumerate((wq, wk, wv)):
                    g = grad_buffer(w)
                    if g is not None:
                        wgrad(dqkv[:, i * D:(i + 1) * D], x, g)
            fuse_bias = need_dx and w_all is not None and gb_all is not None  # the dx dgrad below reads all of dqkv
            if gb_all is not None:
                if not fuse_bias:
                    ops.colsum_accumulate(dqkv, gb_all)
            else:
                for i, b in enumerate((bq, bk, bv)):
                    g = grad_buffer(b)
                    if g is not None:
                        ops.colsum_accumulate(dqkv[:, i * D:(i + 1) * D], g)
            if need_dx:
                dx = torch.empty(M, D, device=dev, dtype=torch.float32)
                res = None if ctx.tail else dy  # residual branch: all rows, or (tail mode) only the last-token rows, added below
                if w_all is not None:
                    ops.gemm(dqkv, w_all, dx, b_mn=True, residual=res, colsum_a=gb_all if fuse_bias else None)
                else:
                    ops.gemm(dqkv[:, :D], wq, dx, b_mn=True, residual=res)
                    ops.gemm(dqkv[:, D:2 * D], wk, dx, b_mn=True, residual=dx)
                    ops.gemm(dqkv[:, 2 * D:], wv, dx, b_mn=True, residual=dx)
                if ctx.tail:
                    dx.view(B, Lq, D)[:, Lq - 1, :].add_(dy)
        else:
            qb, kvb = proj
            dq, dkv = torch.empty_like(qb), torch.empty_like(kvb)
            q = (qb, Lq * D, D)
            k = (kvb, Lk * 2 * D, 2 * D)
            v = (kvb[:, D:], Lk * 2 * D, 2 * D)
            ops.attention_bwd(q, k, v, B, H, Lq, Lk, dh, meta["mode"], meta["layout"], meta["U"], meta["u"], top, dcontext,
                              dq, dkv, dkv[:, D:], dropout=attn_drop)
            g = grad_buffer(wq)
            if g is not None:
                wgrad(dq, x, g)
            g = grad_buffer(bq)
            if g is not None:
                ops.colsum_accumulate(dq, g)
            gw_kv, gb_kv = _fused_grads(wk, wv), _fused_grads(bk, bv)
            if gw_kv is not None:
                wgrad(dkv, cross, gw_kv)
            else:
                for i, w in enumerate((wk, wv)):
                    g = grad_buffer(w)
                    if g is not None:
                        wgrad(dkv[:, i * D:(i + 1) * D], cross, g)
            if gb_kv is not None:
                ops.colsum_accumulate(dkv, gb_kv)
            else:
                for i, b in enumerate((bk, bv)):
                    g = grad_buffer(b)
                    if g is not None:
                        ops.colsum_accumulate(dkv[:, i * D:(i + 1) * D], g)
            if need_dx:
                dx = torch.empty(M, D, device=dev, dtype=torch.float32)
                ops.gemm(dq, wq, dx, b_mn=True, residual=dy)
            if ctx.needs_input_grad[1]:
                dcross = torch.empty_like(cross)
                w_kv = _fused(wk, wv)
                if w_kv is not None:
                    ops.gemm(dkv, w_kv, dcross, b_mn=True)
                else:
                    ops.gemm(dkv[:, :D], wk, dcross, b_mn=True)
                    ops.gemm(dkv[:, D:], wv, dcross, b_mn=True, residual=dcross)
        return (dx, dcross) + (None,) * 10


class FFNBlock(Function):
    """y = x + W2 act(W1 x + b1) + b2   (1x1 Conv1d pair; cross_modal_transformer.py:296-301, TransformerEncoderDecoder.py:48-52)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act, drop=None):
        """drop = {"p", "hidden": (seed, off), "out": (seed, off)}: y = x + dropout(W2 dropout(act(W1 x + b1)) + b2)."""
        M, D = x.shape
        dff = w1.shape[0]
        w1m, w2m = w1.view(dff, D), w2.view(D, dff)
        h = torch.empty(M, dff, device=x.device, dtype=torch.float32)
        # GELU: the forward epilogue stores gelu'(pre) (it shares erf/exp with gelu), the backward only multiplies by it
        pre = torch.empty(M, dff, device=x.device, dtype=torch.float32) if act == ops.ACT_GELU else None
        ops.gemm(x, w1m, h, bias=b1, act=ops.ACT_GELU_SAVE_GRAD if act == ops.ACT_GELU else act, preact=pre)
        y = torch.empty(M, D, device=x.device, dtype=torch.float32)
        if drop:
            ops.dropout(h, h, drop["p"], *drop["hidden"])  # in place: the backward needs the dropped activations for dW2
            ops.gemm(h, w2m, y, bias=b2)
            ops.dropout(y, y, drop["p"], *drop["out"], residual=x)
        else:
            ops.gemm(h, w2m, y, bias=b2, residual=x)
        ctx.drop = drop
        ctx.act = act
        ctx.params = (w1, b1, w2, b2)
        ctx.save_for_backward(x, h, pre)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, h, pre = ctx.saved_tensors
        w1, b1, w2, b2 = ctx.params
        M, D = x.shape
        dff = w1.shape[0]
        dy = dy.contiguous()
        drop = ctx.drop
        dt = dy  # gradient of the conv2 output; dy itself stays the gradient of the residual branch
        if drop:
            dt = ops.dropout(dy, torch.empty_like(dy), drop["p"], *drop["out"])
        g = grad_buffer(w2)
        if g is not None:
            wgrad(dt, h, g.view(D, dff))
        dpre = torch.empty(M, dff, device=x.device, dtype=torch.float32)
        gb2 = grad_buffer(b2)  # bias gradients ride on the dgrad GEMMs that read the same tensor as their A operand
        if ctx.act == ops.ACT_GELU:
            ops.gemm(dt, w2.view(D, dff), dpre, b_mn=True, dact=ops.DACT_SAVED, dact_aux=pre, colsum_a=gb2)
        else:
            # relu'(pre) = [h > 0]; with dropout h holds the DROPPED activations: a dropped unit reads as inactive, and its
            # gradient is zeroed by the mask below anyway
            ops.gemm(dt, w2.view(D, dff), dpre, b_mn=True, dact=ctx.act, dact_aux=h, colsum_a=gb2)
        if drop:
            ops.dropout(dpre, dpre, drop["p"], *drop["hidden"])
        g = grad_buffer(w1)
        if g is not None:
            wgrad(dpre, x, g.view(dff, D))
        gb1 = grad_buffer(b1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, D, device=x.device, dtype=torch.float32)
            ops.gemm(dpre, w1.view(dff, D), dx, b_mn=True, residual=dy, colsum_a=gb1)
        elif gb1 is not None:
            ops.colsum_accumulate(dpre, gb1)
        return dx, None, None, None, None, None, None


class LayerNorm(Function):
    """aten::native_layer_norm over the last dim of a 2-D (possibly row-strided) view, eps = 1e-5."""

    @staticmethod
    def forward(ctx, x, gamma, beta):
        M, D = x.shape
        y = torch.empty(M, D, device=x.device, dtype=torch.float32)
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty(M, device=x.device, dtype=torch.float32)
        ops.layernorm_fwd(x, gamma, beta, y, mean, rstd)
        ctx.params = (gamma, beta)
        ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        gamma, beta = ctx.params
        dy = dy.contiguous()
        dx = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx, grad_buffer(gamma), grad_buffer(beta))
        return dx, None, None


class Linear(Function):
    """y = x W^T + b on a 2-D (possibly row-strided) view (projection heads)."""

    @staticmethod
    def forward(ctx, x, w, b):
        y = torch.empty(x.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
        ops.gemm(x, w, y, bias=b)
        ctx.params = (w, b)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        w, b = ctx.params
        dy = dy.contiguous()
        N = w.shape[0]
        if N % 4:  # TMA pitch: pad the gradient to a multiple of 4 columns (c_out = 66 -> 68)
            pad = torch.zeros(dy.shape[0], (N + 3) // 4 * 4, device=dy.device, dtype=torch.float32)
            pad[:, :N] = dy
            dy_p = pad[:, :N]
        else:
            dy_p = dy
        g = grad_buffer(w)
        if g is not None:
            xc = x if x.stride(0) % 4 == 0 else x.contiguous()
            wgrad(dy_p, xc, g)
        g = grad_buffer(b)
        if g is not None:
            ops.colsum_accumulate(dy, g)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x.shape[0], w.shape[1], device=x.device, dtype=torch.float32)
            ops.gemm(dy_p, w, dx, b_mn=True)
        return dx, None, None


class CircularConv3(Function):
    """Circular Conv1d(k=3) + bias + positional table + time feature, as one GEMM (N = 3D) and a shift-add.

    x2 [n*L, Cp] with Cp >= C a multiple of 4 (columns >= C are zero).  weight [D, C, 3] in the reference layout.
    Returns [n*(L+2*pad-2), D].
    """

    @staticmethod
    def forward(ctx, x2, weight, bias, pe, wtime, n, L, pad):
        D, C, _ = weight.shape
        Cp = x2.shape[1]
        dev = x2.device
        wcat = torch.empty(3 * D, Cp, device=dev, dtype=torch.float32)
        ops.conv3_pack_weight(weight, wcat)
        z = torch.empty(n * L, 3 * D, device=dev, dtype=torch.float32)
        if x2.dtype == torch.bfloat16:  # bf16 operand mode, inference only: the patch embedding stored its features in bf16
            if torch.is_grad_enabled():
                raise RuntimeError("bf16 features reach the token convolution only under torch.no_grad() (ops.bf16_operands)")
            ops.gemm(x2, wcat.to(torch.bfloat16), z)
        else:
            ops.gemm(x2, wcat, z)
        L_out = L + 2 * pad - 2
        y = torch.empty(n * L_out, D, device=dev, dtype=torch.float32)
        ops.conv3_assemble_fwd(z, y, n, L, D, pad, bias=bias, pe=pe, wtime=None if wtime is None else wtime.view(-1))
        ctx.dims = (n, L, pad, D, C, Cp)
        ctx.params = (weight, bias, wtime)
        ctx.save_for_backward(x2, wcat)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, L, pad, D, C, Cp = ctx.dims
        weight, bias, wtime = ctx.params
        x2, wcat = ctx.saved_tensors
        dev = x2.device
        dy = dy.contiguous()
        dz = torch.empty(n * L, 3 * D, device=dev, dtype=torch.float32)
        gt = grad_buffer(wtime)
        ops.conv3_assemble_bwd(dy, dz, n, L, D, pad, dbias=grad_buffer(bias), dwtime=None if gt is None else gt.view(-1))
        g = grad_buffer(weight)
        if g is not None:
            dwcat = torch.zeros(3 * D, Cp, device=dev, dtype=torch.float32)
            ops.gemm(dz, x2, dwcat, a_mn=True, b_mn=True, accumulate=True)
            ops.conv3_unpack_grad(dwcat, g)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(n * L, Cp, device=dev, dtype=torch.float32)
            ops.gemm(dz, wcat, dx, b_mn=True)
        return dx, None, None, None, None, None, None, None


class Dropout(Function):
    """nn.Dropout in training mode on a 2-D view (Informer DataEmbedding, Embedding.py:122-126): stateless Philox mask keyed by
    the site triple (seed, offset, device step base); the backward pass regenerates the same mask."""

    @staticmethod
    def forward(ctx, x, p, site):
        ctx.cfg = (p, site)
        return ops.dropout(x, torch.empty_like(x), p, *site)

    @staticmethod
    def backward(ctx, dy):
        p, site = ctx.cfg
        dy = dy.contiguous()
        return ops.dropout(dy, torch.empty_like(dy), p, *site), None, None


class PatchEmbed(Function):
    """tokens[n*(GG+1), C] = patch rows (GG per frame) x W^T + b, plus the constant -1 token row per frame -- with a gradient for
    the projection (train_backbone: the reference un-freezes its backbone after epoch 10, TimmBackbone.py:123).
    The patches (the FoV-crop output) carry no gradient: the crop has no trainable input."""

    @staticmethod
    def forward(ctx, patches, w, b, GG, round_f16):
        C = w.shape[0]
        n = patches.shape[0] // GG
        tokens = torch.empty(n * (GG + 1), C, device=patches.device, dtype=torch.float32)
        tokens.view(n, GG + 1, C)[:, GG, :] = -1.0
        ops.gemm(patches, w.view(C, -1), tokens, bias=b, out_group=(GG, GG + 1, 0), round_f16=round_f16)
        ctx.dims = (n, GG, C)
        ctx.params = (w, b)
        ctx.save_for_backward(patches)
        return tokens

    @staticmethod
    def backward(ctx, dtokens):
        (patches,) = ctx.saved_tensors
        n, GG, C = ctx.dims
        w, b = ctx.params
        # gradient rows of the patch tokens only (the -1 token is a constant); the fp16 rounding of the features is a straight-through
        dy = dtokens.view(n, GG + 1, C)[:, :GG, :].reshape(n * GG, C)
        g = grad_buffer(w)
        if g is not None:
            wgrad(dy, patches, g.view(C, -1))
        gb = grad_buffer(b)
        if gb is not None:
            ops.colsum_accumulate(dy, gb)
        return None, None, None, None, None


class DistilTail(Function):
    """BatchNorm1d (batch statistics when training; running statistics updated in place) -> ELU -> MaxPool1d(3,2,1)."""

    @staticmethod
    def forward(ctx, z2, gamma, beta, running_mean, running_var, B, Lz, training):
        D = z2.shape[1]
        dev = z2.device
        Lp = (Lz - 1) // 2 + 1
        mean = torch.empty(D, device=dev, dtype=torch.float32)
        rstd = torch.empty(D, device=dev, dtype=torch.float32)
        out = torch.empty(B * Lp, D, device=dev, dtype=torch.float32)
        argmax = torch.empty(B * Lp, D, device=dev, dtype=torch.int8)
        ops.distil_fwd(z2, B, Lz, D, gamma, beta, running_mean, running_var, training, mean, rstd, out, argmax)
        ctx.dims = (B, Lz, D, training)
        ctx.params = (gamma, beta)
        ctx.save_for_backward(z2, mean, rstd, argmax)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, Lz, D, training = ctx.dims
        gamma, beta = ctx.params
        z2, mean, rstd, argmax = ctx.saved_tensors
        dz = torch.empty_like(z2)
        scratch = torch.empty(2 * D, device=z2.device, dtype=torch.float32)
        ops.distil_bwd(z2, B, Lz, D, gamma, beta, mean, rstd, training, argmax, dout.contiguous(), dz, grad_buffer(gamma), grad_buffer(beta),
                       scratch)
        return dz, None, None, None, None, None, None, None


class MotionFeatures(Function):
    """gps [B,T,2] (+ visual [B,T,E]) -> Informer input [B,T,ld] and the per-clip origin angle (routeformer.py:209-235,279-292)."""

    @staticmethod
    def forward(ctx, gps, visual, E, ld, rotate, normalize, mean, std, zero_visual, input_is_motion):
        B, T, _ = gps.shape
        x = torch.empty(B, T, ld, device=gps.device, dtype=torch.float32)
        origin = torch.empty(B, device=gps.device, dtype=torch.float32)
        ops.motion_features(gps, None if zero_visual else visual, x, origin, E, rotate, normalize, mean, std, input_is_motion)
        ctx.E = E
        ctx.zero_visual = zero_visual
        ctx.mark_non_differentiable(origin)
        return x, origin

    @staticmethod
    def backward(ctx, dx, _dorigin):
        dvis = None
        if ctx.needs_input_grad[1] and not ctx.zero_visual:
            dvis = dx[:, :, 5:5 + ctx.E].contiguous()
        return None, dvis, None, None, None, None, None, None, None, None


class DecoderInput(Function):
    """Informer decoder input: history followed by P copies of the last step ("smart") or zeros (Informer.py:125-149)."""

    @staticmethod
    def forward(ctx, x, P, smart):
        B, T, ld = x.shape
        xdec = torch.empty(B, T + P, ld, device=x.device, dtype=torch.float32)
        ops.decoder_input_fwd(x, xdec, P, smart)
        ctx.dims = (B, T, ld, P, smart)
        return xdec

    @staticmethod
    def backward(ctx, dxdec):
        B, T, ld, P, smart = ctx.dims
        dx = torch.zeros(B, T, ld, device=dxdec.device, dtype=torch.float32)
        ops.decoder_input_bwd(dxdec.contiguous(), dx, P, smart)
        return dx, None, None


class TokenStreams(Function):
    """Fused-token buffer [B, n_streams*T, E]: scene / gaze streams + view embeddings + the output-token stream
    (routeformer.py:334-345,443-459).  streams: list of (src or None, dense, F, first, step, emb)."""

    @staticmethod
    def forward(ctx, meta, *tensors):
        B, T, E = meta["B"], meta["T"], meta["E"]
        kinds = meta["streams"]  # [(has_src, dense, F, first, step)]
        n = len(kinds)
        srcs, embs = tensors[:n], tensors[n:]
        dev = embs[0].device
        tokens = torch.empty(B, n * T, E, device=dev, dtype=torch.float32)
        for s, (has_src, dense, F_, first, step) in enumerate(kinds):
            ops.stream_tokens_fwd(srcs[s] if has_src else None, F_, first, step, dense, embs[s].view(-1), tokens, B, T, E, n * T, s * T)
        ctx.meta = meta
        ctx.embs = embs
        return tokens

    @staticmethod
    def backward(ctx, dtokens):
        meta = ctx.meta
        B, T, E = meta["B"], meta["T"], meta["E"]
        kinds = meta["streams"]
        n = len(kinds)
        dtokens = dtokens.contiguous()
        dsrcs = []
        for s, (has_src, dense, F_, first, step) in enumerate(kinds):
            dsrc = None
            if has_src and ctx.needs_input_grad[1 + s]:
                dsrc = torch.empty(B, T if dense else F_, E, device=dtokens.device, dtype=torch.float32)
            g = grad_buffer(ctx.embs[s])
            ops.stream_tokens_bwd(dtokens, dsrc, F_, first, step, dense, None if g is None else g.view(-1), B, T, E, n * T, s * T)
            dsrcs.append(dsrc)
        return (None, *dsrcs, *([None] * n))


class DecodeWaypoints(Function):
    """out[:, :, :2] -> (rotate back) -> (de-normalise) -> last_gps + cumsum (routeformer.py:246-252,350-374)."""

    @staticmethod
    def forward(ctx, out, origin, last_gps, rotate, normalize, mean, std):
        B, P, ld = out.shape
        wp = torch.empty(B, P, 2, device=out.device, dtype=torch.float32)
        motion = torch.empty(B, P, 2, device=out.device, dtype=torch.float32)
        ops.decode_waypoints_fwd(out, origin, last_gps, wp, motion, rotate, normalize, mean, std)
        ctx.cfg = (B, P, ld, rotate, normalize, std)
        ctx.save_for_backward(origin)
        ctx.mark_non_differentiable(motion)
        return wp, motion

    @staticmethod
    def backward(ctx, dwp, _dmotion):
        B, P, ld, rotate, normalize, std = ctx.cfg
        (origin,) = ctx.saved_tensors
        dout = torch.zeros(B, P, ld, device=dwp.device, dtype=torch.float32)
        ops.decode_waypoints_bwd(dwp.contiguous(), origin, dout, rotate, normalize, std)
        return dout, None, None, None, None, None, None


class DiscountedLoss(Function):
    """FutureDiscountedLoss (losses/future_discounted_mse.py:56-95)."""

    @staticmethod
    def forward(ctx, pred, truth, gamma, epsilon, kind):
        pred = pred.contiguous()
        truth = truth.contiguous()
        loss = ops.discounted_loss_fwd(pred, truth, gamma, epsilon, kind)
        ctx.cfg = (gamma, epsilon, kind)
        ctx.save_for_backward(pred, truth)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        gamma, epsilon, kind = ctx.cfg
        pred, truth = ctx.saved_tensors
        dpred = torch.empty_like(pred)
        ops.discounted_loss_bwd(pred, truth, gamma, epsilon, kind, dloss.contiguous().view(1), 1.0, dpred)
        return dpred, None, None, None, None
