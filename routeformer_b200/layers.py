"""Perceive encoder / decoder and their building blocks on the CUDA library.

Class names, constructor signatures, sub-module attribute names (= state_dict keys) and construction order
(= RNG consumption at init) follow routeformer/models/cross_modal_transformer.py, so these classes are
drop-ins for the reference ones; the arithmetic runs in fused blocks of `functional.py`.
Activations travel as row-major 2-D tensors [n*L, D].
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as tF

from . import functional as Fn
from . import ops


def sparse_budget(length: int, factor: int) -> int:
    """u / U of ProbSparse attention: factor * ceil(ln L), capped at L (cross_modal_transformer.py:149-153)."""
    b = factor * int(math.ceil(math.log(length)))
    return b if b < length else length


# ------------------------------------------------------------------------------------------------
# sampled-key index sources.  The reference draws torch.randint(L_K, (L_Q, U)) on the CPU default generator
# inside every ProbAttention.forward (cross_modal_transformer.py:95); reproducing that stream (same calls,
# same order) is what makes outputs comparable at all, so the host keeps drawing on the CPU generator and
# only the resulting int32 tables go to the device.
# ------------------------------------------------------------------------------------------------
class LiveIndexSource:
    """Draws on demand, one small H2D copy per attention call (stand-alone module use)."""

    def __init__(self, device):
        self.device = device
        self.log: List[Tuple[int, int, int]] = []

    def take(self, L_K: int, L_Q: int, U: int):
        idx = torch.randint(L_K, (L_Q, U))
        self.log.append((L_K, L_Q, U))
        return idx.to(torch.int32).unsqueeze(0).to(self.device, non_blocking=True), 0

    def forced_top(self, name: str = ""):
        return None


class PlannedIndexSource:
    """Replays tables that were drawn up-front in reference order and uploaded with ONE copy (Routeformer.forward)."""

    def __init__(self, entries, forced_tops=None):
        """forced_tops (test hook): a list consumed in call order, or {module path: [top, ...]} consumed per path -- the
        kernels then use these top-u selections instead of their own (ProbSparse selection is a discontinuous function of the
        scores; forcing the reference's selections isolates the arithmetic from tie-breaking)."""
        self.entries = list(entries)  # [(key, idx_dev [g,L_Q,U] int32, idx_group)]
        self.pos = 0
        self.tops = forced_tops if (forced_tops is None or isinstance(forced_tops, dict)) else list(forced_tops)
        self.top_pos = 0

    def take(self, L_K: int, L_Q: int, U: int):
        key, idx, group = self.entries[self.pos]
        self.pos += 1
        if key != (L_K, L_Q, U):
            raise RuntimeError(f"index plan out of order: planned {key}, requested {(L_K, L_Q, U)}")
        return idx, group

    def forced_top(self, name: str = ""):
        if self.tops is None:
            return None
        if isinstance(self.tops, dict):
            return self.tops[name].pop(0)
        t = self.tops[self.top_pos]
        self.top_pos += 1
        return t


def _site(device, p: float, where: str, rows: int, cols: int):
    """One dropout call site of this step: (seed, offset) of its Philox sub-stream."""
    return ops.DropoutStream.next(device, where, rows, cols)


# ------------------------------------------------------------------------------------------------
# parameter containers (reference module tree)
# ------------------------------------------------------------------------------------------------
class PositionalEmbedding(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, d_model).float()
        position = torch.arange(0, max_len).float().unsqueeze(1)
        div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))

    def table(self) -> torch.Tensor:
        return self.pe[0]


class TokenEmbedding(nn.Module):
    def __init__(self, c_in, d_model, bias=True):
        super().__init__()
        self.tokenConv = nn.Conv1d(c_in, d_model, kernel_size=3, padding=1, padding_mode="circular", bias=bias)
        nn.init.kaiming_normal_(self.tokenConv.weight, mode="fan_in", nonlinearity="leaky_relu")


class _InnerAttention(nn.Module):
    """Parameter-free placeholder keeping the reference attribute name `inner_attention`."""

    def __init__(self, kind: str, factor: int):
        super().__init__()
        self.kind, self.factor = kind, factor


class AttentionLayer(nn.Module):
    """q/k/v/out projections around ProbSparse or full attention; forward includes the caller's residual add."""

    def __init__(self, kind: str, factor: int, d_model: int, n_heads: int, informer_layout: bool = False):
        super().__init__()
        self.inner_attention = _InnerAttention(kind, factor)
        self.query_projection = nn.Linear(d_model, d_model)
        self.key_projection = nn.Linear(d_model, d_model)
        self.value_projection = nn.Linear(d_model, d_model)
        self.out_projection = nn.Linear(d_model, d_model)
        self.n_heads = n_heads
        self.informer_layout = informer_layout

    def block(self, x2, cross2, B, Lq, Lk, draw, record=None, name="", p_drop: float = 0.0, tail: bool = False):
        """p_drop > 0 (training): `x + dropout(attention(x))` of the calling layer and, for full attention, the dropout on the
        softmax probabilities (cross_modal_transformer.py:63,224,227,295)."""
        kind, factor = self.inner_attention.kind, self.inner_attention.factor
        mode = {"prob": ops.ATTN_PROB, "prob_masked": ops.ATTN_PROB_MASKED, "full": ops.ATTN_FULL}[kind]
        meta = dict(B=B, H=self.n_heads, Lq=Lq, Lk=Lk, mode=mode,
                    layout=ops.LAYOUT_BHLD if self.informer_layout else ops.LAYOUT_BLHD,
                    U=0, u=0, idx_group=0, record=record, name=name, tail=tail and cross2 is None and not self.informer_layout)
        idx = None
        if kind != "full":
            meta["U"], meta["u"] = sparse_budget(Lk, factor), sparse_budget(Lq, factor)
            idx, meta["idx_group"] = draw.take(Lk, Lq, meta["U"])
            meta["forced_top"] = draw.forced_top(name)
        if p_drop > 0.0:
            dev, D = x2.device, self.out_projection.weight.shape[0]
            prob = _site(dev, p_drop, name + ".prob", B * self.n_heads * Lq, Lk) if kind == "full" else None
            meta["drop"] = {"p": p_drop, "prob": prob, "out": _site(dev, p_drop, name + ".out", x2.shape[0], D)}
        q, k, v, o = self.query_projection, self.key_projection, self.value_projection, self.out_projection
        return Fn.AttentionBlock.apply(x2, cross2, q.weight, q.bias, k.weight, k.bias, v.weight, v.bias, o.weight, o.bias, idx, meta)


class EncoderLayer(nn.Module):
    def __init__(self, attention: AttentionLayer, d_model, d_ff=None, dropout=0.1, activation="relu"):
        super().__init__()
        d_ff = d_ff or 4 * d_model
        self.attention = attention
        self.conv1 = nn.Conv1d(d_model, d_ff, kernel_size=1)
        self.conv2 = nn.Conv1d(d_ff, d_model, kernel_size=1)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.p_drop = dropout
        self.act = ops.ACT_RELU if activation == "relu" else ops.ACT_GELU

    def _ffn_drop(self, x2, name):
        if not (self.training and self.p_drop > 0.0):
            return None
        M, dff, D = x2.shape[0], self.conv1.weight.shape[0], self.conv2.weight.shape[0]
        return {"p": self.p_drop, "hidden": _site(x2.device, self.p_drop, name + ".ffn_hidden", M, dff),
                "out": _site(x2.device, self.p_drop, name + ".ffn_out", M, D)}

    def run(self, x2, B, L, draw, record=None, name="", tail: bool = False):
        """tail=True (last layer of an encoder whose caller keeps only the last token): returns [B, D] instead of [B*L, D]."""
        p = self.p_drop if self.training else 0.0
        x2 = self.attention.block(x2, None, B, L, L, draw, record, name + ".attention", p, tail=tail and p == 0.0)
        x2 = Fn.LayerNorm.apply(x2, self.norm1.weight, self.norm1.bias)
        x2 = Fn.FFNBlock.apply(x2, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.act, self._ffn_drop(x2, name))
        return Fn.LayerNorm.apply(x2, self.norm2.weight, self.norm2.bias)


class DecoderLayer(nn.Module):
    def __init__(self, self_attention: AttentionLayer, cross_attention: AttentionLayer, d_model, d_ff=None, dropout=0.1,
                 activation="relu"):
        super().__init__()
        d_ff = d_ff or 4 * d_model
        self.self_attention = self_attention
        self.cross_attention = cross_attention
        self.conv1 = nn.Conv1d(d_model, d_ff, kernel_size=1)
        self.conv2 = nn.Conv1d(d_ff, d_model, kernel_size=1)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.p_drop = dropout
        self.act = ops.ACT_RELU if activation == "relu" else ops.ACT_GELU

    _ffn_drop = EncoderLayer._ffn_drop

    def run(self, x2, cross2, B, L, S, draw, record=None, name=""):
        p = self.p_drop if self.training else 0.0
        x2 = self.self_attention.block(x2, None, B, L, L, draw, record, name + ".self_attention", p)
        x2 = Fn.LayerNorm.apply(x2, self.norm1.weight, self.norm1.bias)
        x2 = self.cross_attention.block(x2, cross2, B, L, S, draw, record, name + ".cross_attention", p)
        x2 = Fn.LayerNorm.apply(x2, self.norm2.weight, self.norm2.bias)
        x2 = Fn.FFNBlock.apply(x2, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.act, self._ffn_drop(x2, name))
        return Fn.LayerNorm.apply(x2, self.norm3.weight, self.norm3.bias)


class Encoder(nn.Module):
    def __init__(self, attn_layers, conv_layers=None, norm_layer=None):
        super().__init__()
        self.attn_layers = nn.ModuleList(attn_layers)
        self.conv_layers = nn.ModuleList(conv_layers) if conv_layers is not None else None
        self.norm = norm_layer


class Decoder(nn.Module):
    def __init__(self, layers, norm_layer=None, projection=None):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.norm = norm_layer
        if projection is not None:
            self.projection = projection


def _pad_channels(x2: torch.Tensor) -> torch.Tensor:
    """TMA needs a 16 B row pitch: zero-pad the channel dim to a multiple of 4 (only the 2-channel gaze input needs it)."""
    C = x2.shape[1]
    return x2 if C % 4 == 0 else tF.pad(x2, (0, 4 - C % 4))


def _select_tail(h2: torch.Tensor, n: int, L: int, out_len: int) -> torch.Tensor:
    """Rows of the last `out_len` tokens of every sequence, as a 2-D view/copy (cross_modal_transformer.py:433)."""
    if out_len >= L:
        return h2
    D = h2.shape[1]
    h3 = h2.view(n, L, D)
    if out_len == 1:
        return h3[:, L - 1, :]  # strided 2-D view, consumed in place by the LN / GEMM kernels
    return h3[:, L - out_len:, :].reshape(n * out_len, D)


class PerceiveEncoder(nn.Module):
    """Drop-in for cross_modal_transformer.PerceiveEncoder (:372-433)."""

    def __init__(self, in_channels, out_channels, out_len, factor=5, d_model=128, n_heads=8, layers=3, d_ff=None, dropout=0.1,
                 activation="gelu", output_attention=False):
        super().__init__()
        if output_attention:
            raise NotImplementedError("output_attention=True is not supported by the fused attention kernel")
        self.pred_len = out_len
        self.d_model = d_model
        d_ff = d_ff or 4 * d_model
        self.value_embedding = TokenEmbedding(in_channels, d_model)
        self.position_embedding = PositionalEmbedding(d_model)
        self.encoder = Encoder(
            [EncoderLayer(AttentionLayer("prob", factor, d_model, n_heads), d_model, d_ff, dropout=dropout, activation=activation)
             for _ in range(layers)],
            None, norm_layer=nn.LayerNorm(d_model))
        self.projection = nn.Linear(d_model, out_channels, bias=True)

    def encode(self, x2: torch.Tensor, n: int, L: int, draw, record=None, name="", samples: int = 1) -> torch.Tensor:
        """x2 [n*L, Cp] (channel-padded) -> [samples*n*min(L,out_len), out_channels].  samples > 1 (Routeformer.forward_samples):
        the deterministic token embedding is computed once, the layers run on `samples` copies of the sequences, sample-major."""
        conv = self.value_embedding.tokenConv
        h = Fn.CircularConv3.apply(x2, conv.weight, conv.bias, self.position_embedding.table(), None, n, L, 1)
        if samples > 1:
            h = h.repeat(samples, 1)
            n = n * samples
        n_layers = len(self.encoder.attn_layers)
        tail = False
        for i, layer in enumerate(self.encoder.attn_layers):
            # the frame encoder keeps one token of 65: its last layer computes LayerNorms / FFN / out-projection for that token only
            tail = i == n_layers - 1 and self.pred_len == 1 and L > 1 and not (layer.training and layer.p_drop > 0.0)
            h = layer.run(h, n, L, draw, record, f"{name}.encoder.attn_layers.{i}", tail=tail)
        if not tail:
            h = _select_tail(h, n, L, self.pred_len)
        h = Fn.LayerNorm.apply(h, self.encoder.norm.weight, self.encoder.norm.bias)
        return Fn.Linear.apply(h, self.projection.weight, self.projection.bias)

    def forward(self, x_enc: torch.Tensor, draw=None) -> torch.Tensor:
        n, L, C = x_enc.shape
        draw = draw or LiveIndexSource(x_enc.device)
        if self.training and any(layer.p_drop > 0.0 for layer in self.encoder.attn_layers):
            ops.DropoutStream.begin_step(x_enc.device)
        x2 = _pad_channels(x_enc.to(torch.float32).reshape(n * L, C))
        out = self.encode(x2, n, L, draw)
        return out.view(n, min(L, self.pred_len), -1)


class PerceiveDecoder(nn.Module):
    """Drop-in for cross_modal_transformer.PerceiveDecoder (:436-503); `mix` must be False as in routeformer.py:114."""

    def __init__(self, query_channels, value_channels, out_channels, out_len, factor=5, n_heads=8, layers=2, d_ff=None, dropout=0.1,
                 activation="gelu", mix=True):
        super().__init__()
        if mix:
            raise NotImplementedError("PerceiveDecoder(mix=True) is not on the Routeformer path (routeformer.py:114 passes mix=False)")
        self.pred_len = out_len
        d_model = value_channels
        d_ff = d_ff or 4 * d_model
        self.value_embedding = TokenEmbedding(query_channels, d_model)
        self.position_embedding = PositionalEmbedding(d_model)
        self.decoder = Decoder(
            [DecoderLayer(AttentionLayer("prob_masked", factor, d_model, n_heads), AttentionLayer("full", factor, d_model, n_heads),
                          d_model, d_ff, dropout=dropout, activation=activation) for _ in range(layers)],
            norm_layer=nn.LayerNorm(d_model))
        self.projection = nn.Linear(d_model, out_channels, bias=True)

    def decode(self, enc2, dec2, n, S, L, draw, record=None, name=""):
        conv = self.value_embedding.tokenConv
        h = Fn.CircularConv3.apply(dec2, conv.weight, conv.bias, self.position_embedding.table(), None, n, L, 1)
        for i, layer in enumerate(self.decoder.layers):
            h = layer.run(h, enc2, n, L, S, draw, record, f"{name}.decoder.layers.{i}")
        h = _select_tail(h, n, L, self.pred_len)
        h = Fn.LayerNorm.apply(h, self.decoder.norm.weight, self.decoder.norm.bias)
        return Fn.Linear.apply(h, self.projection.weight, self.projection.bias)

    def forward(self, x_enc: torch.Tensor, x_dec: torch.Tensor, draw=None) -> torch.Tensor:
        n, S, Cv = x_enc.shape
        _, L, Cq = x_dec.shape
        draw = draw or LiveIndexSource(x_dec.device)
        if self.training and any(layer.p_drop > 0.0 for layer in self.decoder.layers):
            ops.DropoutStream.begin_step(x_dec.device)
        enc2 = x_enc.to(torch.float32).reshape(n * S, Cv)
        dec2 = _pad_channels(x_dec.to(torch.float32).reshape(n * L, Cq))
        out = self.decode(enc2, dec2, n, S, L, draw)
        return out.view(n, min(L, self.pred_len), -1)
