"""TEST INFRASTRUCTURE ONLY -- functional CPU restatement of the Routeformer hot path.

Every function below restates, in plain functional PyTorch over a ``state_dict`` (no
nn.Module, no Lightning), what one reference function computes, and cites the reference
file:line it follows (paths relative to /root/reference).  Autograd works through all of
it, so gradient parity is checked with ``torch.autograd.grad`` on these functions.

Parity status: PINNED against the unmodified reference (see oracle/__init__.py).

Conventions
-----------
* ``sd`` is a reference-layout ``state_dict`` (same keys/shapes as the reference model).
* ``draw`` is an index source: ``draw(L_K, L_Q, U) -> int64 [L_Q, U]``.  The default
  (`CpuRandint`) calls ``torch.randint(L_K, (L_Q, U))`` on the CPU default generator in
  exactly the reference's order, so ``torch.manual_seed(s)`` reproduces the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# configuration (only the fields the hot path reads)
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    # Informer / GPS backbone (gps_backbone/config.py:9-75, full_comparison.py:159-178)
    seq_len: int = 40
    pred_len: int = 30
    d_model: int = 832
    n_heads: int = 8
    e_layers: int = 6
    d_layers: int = 1
    d_ff: int = 3328
    factor: int = 4
    distil: bool = True
    activation: str = "relu"
    # Routeformer (models/config.py:10-107)
    decoder_mode: str = "smart"
    with_video: bool = False
    with_scene: bool = True
    with_gaze: bool = False
    dense_prediction: bool = False
    image_embedding_size: int = 64
    encoder_hidden_size: int = 64
    encoder_heads: int = 8
    encoder_layers: int = 8
    encoder_d_ff: int = 256
    cross_modal_decoder_heads: int = 8
    cross_modal_decoder_layers: int = 2
    output_fps: int = 5
    video_fps: int = 1
    gaze_fps: int = 1
    rotate_motion: bool = False
    normalize_motion: bool = False
    motion_mean: float = 0.0
    motion_std: float = 1.0
    only_motion: bool = False
    view_dropout: float = 0.0
    gaze_dropout: float = 0.0
    autoregressive: bool = False  # models/config.py:36-38 (eval only)
    autoregressive_step_size: int = 1
    # Perceive* hard defaults (cross_modal_transformer.py:378-379,443)
    perceive_factor: int = 5
    perceive_d_model: int = 128

    @property
    def enc_in(self) -> int:  # gps_backbone/config.py:57-70
        return 5 + (self.encoder_hidden_size if self.with_video else 0)

    @property
    def c_out(self) -> int:  # gps_backbone/config.py:42-52
        return self.enc_in - 3 if self.dense_prediction else 2


@dataclass
class BackboneSpec:
    """Build-defined random-init visual encoder (replaces the un-vendored timm model).

    crop/resample (bilinear, zeros padding, align_corners=False) -> per-channel normalise
    -> Conv2d(3, channels, k=patch, s=patch).  ``fov`` selects the sampling window:
    "frame" = zero-pad bottom/right to a square and resize (nearest equivalent of
    video_backbone/TimmBackbone.py:164-177), "gaze" = a window of ``window`` x frame size
    centred on a per-frame fixation.
    """

    image_size: int = 256
    patch: int = 32
    channels: int = 1024
    window: float = 0.5
    mean: Tuple[float, float, float] = (0.485, 0.456, 0.406)
    std: Tuple[float, float, float] = (0.229, 0.224, 0.225)

    @property
    def grid(self) -> int:
        return self.image_size // self.patch


# --------------------------------------------------------------------------------------
# index sources (cross_modal_transformer.py:95, SelfAttentionFamily.py:94)
# --------------------------------------------------------------------------------------
class CpuRandint:
    """``torch.randint(L_K, (L_Q, U))`` on the CPU default generator; logs every call."""

    def __init__(self):
        self.log: List[Tuple[int, int, int]] = []
        self.drawn: List[Tensor] = []

    def __call__(self, L_K: int, L_Q: int, U: int) -> Tensor:
        idx = torch.randint(L_K, (L_Q, U))
        self.log.append((L_K, L_Q, U))
        self.drawn.append(idx)
        return idx


class Replay:
    """Replays a pre-drawn list of index tensors in order (shape-checked)."""

    def __init__(self, tensors: List[Tensor], tops=None):
        self.tensors = list(tensors)
        self.pos = 0
        self.tops = tops if (tops is None or isinstance(tops, dict)) else list(tops)
        self.top_pos = 0

    def next_top(self, where: str = "") -> Optional[Tensor]:
        """Forced top-u selection for the next ProbSparse call (test hook, see prob_attention).

        ``tops`` is either a list consumed in call order or a dict {module path: [tensor, ...]} consumed per path."""
        if self.tops is None:
            return None
        if isinstance(self.tops, dict):
            t = self.tops[where].pop(0)
        else:
            t = self.tops[self.top_pos]
            self.top_pos += 1
        return t.long().cpu()


# --------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------
def pe_table(length: int, d_model: int) -> Tensor:
    """Sinusoidal table, cross_modal_transformer.py:332-349 / Embedding.py:8-26."""
    pos = torch.arange(0, length).float().unsqueeze(1)
    div = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
    pe = torch.zeros(length, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def circular_conv3(x: Tensor, weight: Tensor, bias: Optional[Tensor], pad: int) -> Tensor:
    """Circular Conv1d, kernel 3, over the token axis.

    x [n, L, C], weight [D, C, 3] -> [n, L + 2*pad - 2, D].
    y[t] = sum_j W[:, :, j] x[(t - pad + j) mod L]  (cross_modal_transformer.py:356-368 with
    pad=1; TransformerEncoderDecoder.py:12-18 with pad=2; Embedding.py:32-45 pad=1 no bias).
    Uses the same ATen ops as nn.Conv1d(padding_mode="circular") so results are bit-equal to
    the reference; ``circular_conv3_explicit`` is the index-level statement of the same sum.
    """
    xt = x.permute(0, 2, 1)
    xt = F.pad(xt, (pad, pad), mode="circular")
    return F.conv1d(xt, weight, bias).transpose(1, 2)


def circular_conv3_explicit(x: Tensor, weight: Tensor, bias: Optional[Tensor], pad: int) -> Tensor:
    """Index-level form of `circular_conv3` (what the CUDA path implements)."""
    n, L, C = x.shape
    L_out = L + 2 * pad - 2
    t = torch.arange(L_out)
    y = x.new_zeros(n, L_out, weight.shape[0])
    for j in range(3):
        src = (t - pad + j) % L
        y = y + x[:, src, :] @ weight[:, :, j].t()
    if bias is not None:
        y = y + bias
    return y


def sparse_budget(length: int, factor: int) -> int:
    """u / U of ProbSparse attention (cross_modal_transformer.py:149-153)."""
    b = factor * int(math.ceil(math.log(length)))
    return b if b < length else length


def prob_attention(
    q: Tensor, k: Tensor, v: Tensor, idx: Tensor, factor: int, masked: bool,
    forced_top: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    """ProbSparse attention (cross_modal_transformer.py:88-166; SelfAttentionFamily.py:87-165).

    q [B,Lq,H,E], k [B,Lk,H,E], v [B,Lk,H,Dv], idx int64 [Lq,U] shared by every (b,h).
    Returns (context [B,H,Lq,Dv], top [B,H,u], measure [B,H,Lq]).  The caller applies the
    layout quirk (Perceive*: transpose back to [B,Lq,H,Dv]; Informer: use [B,H,Lq,Dv] as is).

    ``forced_top`` (test hook): top-u selection is a discontinuous function of the scores, so a
    second implementation with different rounding can legitimately pick a different query at a
    near-tie.  Parity tests therefore (i) replay the selection of the implementation under test
    here and compare outputs tightly, and (ii) check that selection against ``measure``: every
    selected query must be within a small gap of this function's own top-u set.
    """
    B, Lq, H, E = q.shape
    Lk = k.shape[1]
    Dv = v.shape[-1]
    Q, K, V = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    U, u = sparse_budget(Lk, factor), sparse_budget(Lq, factor)
    assert tuple(idx.shape) == (Lq, U), (idx.shape, Lq, U)
    # sampled scores: s[b,h,i,j] = Q[i] . K[idx[i,j]]   (:94-97)
    k_sample = K.unsqueeze(-3).expand(B, H, Lq, Lk, E)[:, :, torch.arange(Lq).unsqueeze(1), idx, :]
    s = torch.matmul(Q.unsqueeze(-2), k_sample.transpose(-2, -1)).squeeze(-2)
    # sparsity measure; note the division by Lk, not U (:100)
    measure = s.max(-1)[0] - s.sum(-1) / Lk
    top = measure.topk(u, sorted=False)[1]  # [B,H,u]  (:101)
    if forced_top is not None:
        assert forced_top.shape == top.shape
        top = forced_top.to(top.device)
    q_top = torch.gather(Q, 2, top.unsqueeze(-1).expand(B, H, u, E))
    scores = torch.matmul(q_top, K.transpose(-2, -1)) * (1.0 / math.sqrt(E))  # (:107,:158-160)
    if masked:  # (:113-121, ProbMask :22-33)
        assert Lq == Lk
        ctx = V.cumsum(dim=-2)
        key_pos = torch.arange(Lk, device=q.device)
        scores = scores.masked_fill(key_pos.view(1, 1, 1, Lk) > top.unsqueeze(-1), float("-inf"))
    else:
        ctx = V.mean(dim=-2, keepdim=True).expand(B, H, Lq, Dv)
    attn = torch.softmax(scores, dim=-1)
    ctx = ctx.scatter(2, top.unsqueeze(-1).expand(B, H, u, Dv), torch.matmul(attn, V))  # (:131-133)
    return ctx, top, measure


# Training-mode feature dropout of the Perceive modules (nn.Dropout at cross_modal_transformer.py:63,224-231,295-299).  The masks
# come from a caller-installed hook `fn(x, where) -> dropped x` (tests replay the CUDA path's Philox masks through it); without a
# hook dropout is the identity, i.e. the parity configuration feature_dropout = 0.
_DROPOUT_HOOK: Optional[Callable] = None
_DROPOUT_INFORMER = False  # hook applied at the Informer's dropout sites (GPSBackboneConfig.dropout > 0; paper config: 0.0)
_DROPOUT_PERCEIVE = True   # hook applied at the Perceive modules' sites (feature_dropout > 0)


class dropout_hook:
    def __init__(self, fn: Optional[Callable], informer: bool = False, perceive: bool = True):
        self.fn, self.informer, self.perceive = fn, informer, perceive

    def __enter__(self):
        global _DROPOUT_HOOK, _DROPOUT_INFORMER, _DROPOUT_PERCEIVE
        self.prev, _DROPOUT_HOOK = _DROPOUT_HOOK, self.fn
        self.prev_inf, _DROPOUT_INFORMER = _DROPOUT_INFORMER, self.informer
        self.prev_per, _DROPOUT_PERCEIVE = _DROPOUT_PERCEIVE, self.perceive

    def __exit__(self, *exc):
        global _DROPOUT_HOOK, _DROPOUT_INFORMER, _DROPOUT_PERCEIVE
        _DROPOUT_HOOK = self.prev
        _DROPOUT_INFORMER = self.prev_inf
        _DROPOUT_PERCEIVE = self.prev_per


def _drop(x: Tensor, where: str, enabled: Optional[bool] = None) -> Tensor:
    """enabled: None = a Perceive-module site (follows _DROPOUT_PERCEIVE); otherwise the caller's decision."""
    if enabled is None:
        enabled = _DROPOUT_PERCEIVE
    return _DROPOUT_HOOK(x, where) if (_DROPOUT_HOOK is not None and enabled) else x


def full_attention(q: Tensor, k: Tensor, v: Tensor, where: Optional[str] = None) -> Tensor:
    """Unmasked softmax attention (cross_modal_transformer.py:51-69).  Returns [B,Lq,H,Dv]."""
    E = q.shape[-1]
    scores = torch.einsum("blhe,bshe->bhls", q, k)
    attn = torch.softmax(scores * (1.0 / math.sqrt(E)), dim=-1)
    if where is not None:
        attn = _drop(attn, where + ".prob")  # :63
    return torch.einsum("bhls,bshd->blhd", attn, v).contiguous()


def _lin(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _act(name: str, x: Tensor) -> Tensor:
    return F.relu(x) if name == "relu" else F.gelu(x)


def attention_layer(
    sd: SD,
    p: str,
    xq: Tensor,
    xkv: Tensor,
    n_heads: int,
    kind: str,
    factor: int,
    draw: Callable,
    informer_layout: bool,
    tops: Optional[list] = None,
) -> Tensor:
    """q/k/v projections + inner attention + out projection.

    Perceive*: cross_modal_transformer.py:169-198 (context transposed back to [B,L,H,dh]).
    Informer : SelfAttentionFamily.py:168-194 -- the [B,H,L,dh] context is ``view``-ed as
    [B,L,H*dh] WITHOUT transposing back (:165,:192).
    kind in {"prob", "prob_masked", "full"}.
    """
    B, L, _ = xq.shape
    S = xkv.shape[1]
    q = _lin(sd, p + ".query_projection", xq).view(B, L, n_heads, -1)
    k = _lin(sd, p + ".key_projection", xkv).view(B, S, n_heads, -1)
    v = _lin(sd, p + ".value_projection", xkv).view(B, S, n_heads, -1)
    if kind == "full":
        out = full_attention(q, k, v, None if informer_layout else p)
    else:
        idx = draw(S, L, sparse_budget(S, factor))
        forced = draw.next_top(p) if hasattr(draw, "next_top") else None
        ctx, top, measure = prob_attention(q, k, v, idx, factor, kind == "prob_masked", forced)
        if tops is not None:
            tops.append({"where": p, "top": top, "measure": measure.detach()})
        out = ctx.contiguous() if informer_layout else ctx.transpose(2, 1).contiguous()
    return _lin(sd, p + ".out_projection", out.view(B, L, -1))


def ffn(sd: SD, p: str, x: Tensor, act: str, perceive: bool = False) -> Tensor:
    """1x1 Conv1d -> act -> dropout -> 1x1 Conv1d -> dropout (cross_modal_transformer.py:297-299)."""
    # same ATen ops as the reference (conv1d over the transposed tensor) so the pin is bit-exact;
    # mathematically h = act(x W1^T + b1), y = h W2^T + b2.
    h = _act(act, F.conv1d(x.transpose(-1, 1), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"]))
    h = _drop(h.transpose(-1, 1), p + ".ffn_hidden", perceive).transpose(-1, 1)
    y = F.conv1d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"]).transpose(-1, 1)
    return _drop(y, p + ".ffn_out", perceive)


def encoder_layer(sd, p, x, n_heads, factor, act, draw, informer_layout, tops=None) -> Tensor:
    """Post-norm encoder block (cross_modal_transformer.py:288-301; TransformerEncoderDecoder.py:43-53)."""
    perceive = _DROPOUT_INFORMER if informer_layout else _DROPOUT_PERCEIVE  # TransformerEncoderDecoder.py:46-50: same three sites
    a = attention_layer(sd, p + ".attention", x, x, n_heads, "prob", factor, draw, informer_layout, tops)
    x = _ln(sd, p + ".norm1", x + _drop(a, p + ".attention.out", perceive))
    return _ln(sd, p + ".norm2", x + ffn(sd, p, x, act, perceive))


def decoder_layer(sd, p, x, cross, n_heads, factor, act, draw, informer_layout, cross_kind, tops=None) -> Tensor:
    """Decoder block (cross_modal_transformer.py:223-233; TransformerEncoderDecoder.py:104-116)."""
    perceive = _DROPOUT_INFORMER if informer_layout else _DROPOUT_PERCEIVE  # TransformerEncoderDecoder.py:106-113
    a = attention_layer(sd, p + ".self_attention", x, x, n_heads, "prob_masked", factor, draw, informer_layout, tops)
    x = _ln(sd, p + ".norm1", x + _drop(a, p + ".self_attention.out", perceive))
    c = attention_layer(sd, p + ".cross_attention", x, cross, n_heads, cross_kind, factor, draw, informer_layout, tops)
    x = _ln(sd, p + ".norm2", x + _drop(c, p + ".cross_attention.out", perceive))
    return _ln(sd, p + ".norm3", x + ffn(sd, p, x, act, perceive))


# --------------------------------------------------------------------------------------
# Perceive encoder / decoder
# --------------------------------------------------------------------------------------
def perceive_encoder(sd: SD, p: str, x: Tensor, out_len: int, cfg: OracleConfig, draw, tops=None) -> Tensor:
    """PerceiveEncoder.forward (cross_modal_transformer.py:425-433).  x [n,L,C] -> [n,out_len,out]."""
    n, L, _ = x.shape
    D = sd[p + ".value_embedding.tokenConv.weight"].shape[0]
    h = circular_conv3(x, sd[p + ".value_embedding.tokenConv.weight"], sd[p + ".value_embedding.tokenConv.bias"], 1)
    h = h + sd[p + ".position_embedding.pe"][0, :L] if (p + ".position_embedding.pe") in sd else h + pe_table(L, D)
    n_layers = 1 + max(int(k.split(".attn_layers.")[1].split(".")[0]) for k in sd if k.startswith(p + ".encoder.attn_layers."))
    for i in range(n_layers):
        h = encoder_layer(
            sd, f"{p}.encoder.attn_layers.{i}", h, cfg.encoder_heads, cfg.perceive_factor, "gelu", draw, False, tops
        )
    h = _ln(sd, p + ".encoder.norm", h)
    return _lin(sd, p + ".projection", h)[:, -out_len:, :]


def perceive_decoder(sd: SD, p: str, x_enc: Tensor, x_dec: Tensor, out_len: int, cfg: OracleConfig, draw, tops=None) -> Tensor:
    """PerceiveDecoder.forward (cross_modal_transformer.py:498-503); mix=False (routeformer.py:114)."""
    L = x_dec.shape[1]
    D = sd[p + ".value_embedding.tokenConv.weight"].shape[0]
    h = circular_conv3(x_dec, sd[p + ".value_embedding.tokenConv.weight"], sd[p + ".value_embedding.tokenConv.bias"], 1)
    h = h + (sd[p + ".position_embedding.pe"][0, :L] if (p + ".position_embedding.pe") in sd else pe_table(L, D))
    for i in range(cfg.cross_modal_decoder_layers):
        h = decoder_layer(
            sd, f"{p}.decoder.layers.{i}", h, x_enc, cfg.cross_modal_decoder_heads, cfg.perceive_factor,
            "gelu", draw, False, "full", tops,
        )
    h = _ln(sd, p + ".decoder.norm", h)
    return _lin(sd, p + ".projection", h)[:, -out_len:, :]


# --------------------------------------------------------------------------------------
# Informer
# --------------------------------------------------------------------------------------
def distil_block(sd: SD, p: str, x: Tensor, training: bool, bn_updates: Optional[dict] = None) -> Tensor:
    """ConvLayer (TransformerEncoderDecoder.py:9-29): circ conv k3 pad2 -> BN1d -> ELU -> MaxPool(3,2,1)."""
    z = circular_conv3(x, sd[p + ".downConv.weight"], sd[p + ".downConv.bias"], 2)  # [B, L+2, D]
    zt = z.transpose(1, 2)
    rm, rv = sd[p + ".norm.running_mean"], sd[p + ".norm.running_var"]
    if training:
        rm2, rv2 = rm.clone(), rv.clone()
        zt = F.batch_norm(zt, rm2, rv2, sd[p + ".norm.weight"], sd[p + ".norm.bias"], True, 0.1, 1e-5)
        if bn_updates is not None:
            bn_updates[p + ".norm.running_mean"] = rm2
            bn_updates[p + ".norm.running_var"] = rv2
    else:
        zt = F.batch_norm(zt, rm, rv, sd[p + ".norm.weight"], sd[p + ".norm.bias"], False, 0.1, 1e-5)
    zt = F.max_pool1d(F.elu(zt), kernel_size=3, stride=2, padding=1)
    return zt.transpose(1, 2)


def informer_embedding(sd: SD, p: str, x: Tensor) -> Tensor:
    """DataEmbedding with timeF/freq="m" (Embedding.py:111-126; Informer.py:119-123,150-153)."""
    L = x.shape[1]
    D = sd[p + ".value_embedding.tokenConv.weight"].shape[0]
    t = torch.arange(L, dtype=torch.float32, device=x.device).view(1, L, 1)
    h = circular_conv3(x, sd[p + ".value_embedding.tokenConv.weight"], None, 1)
    h = h + F.linear(t, sd[p + ".temporal_embedding.embed.weight"])
    pe = sd[p + ".position_embedding.pe"][0, :L] if (p + ".position_embedding.pe") in sd else pe_table(L, D)
    return _drop(h + pe, p + ".dropout", _DROPOUT_INFORMER)  # Embedding.py:126


def informer(sd: SD, p: str, x: Tensor, cfg: OracleConfig, draw, training: bool = False,
             pred_len: Optional[int] = None, bn_updates: Optional[dict] = None, tops=None) -> Tensor:
    """Informer.forward (gps_backbone/Informer.py:105-167).  x [B,T,enc_in] -> [B,P,c_out]."""
    P = cfg.pred_len if pred_len is None else pred_len
    if cfg.decoder_mode == "smart":  # (:125-136)
        x_dec = torch.cat([x, x[:, -1:, :].repeat(1, P, 1)], dim=1)
    else:  # (:137-149)
        x_dec = torch.cat([x, torch.zeros(x.shape[0], P, x.shape[2], dtype=torch.float32, device=x.device)], dim=1)
    h = informer_embedding(sd, p + ".enc_embedding", x)
    for i in range(cfg.e_layers):  # TransformerEncoderDecoder.py:62-80
        h = encoder_layer(sd, f"{p}.encoder.attn_layers.{i}", h, cfg.n_heads, cfg.factor, cfg.activation, draw, True, tops)
        if cfg.distil and i < cfg.e_layers - 1:
            h = distil_block(sd, f"{p}.encoder.conv_layers.{i}", h, training, bn_updates)
    enc = _ln(sd, p + ".encoder.norm", h)
    d = informer_embedding(sd, p + ".dec_embedding", x_dec)
    for i in range(cfg.d_layers):
        d = decoder_layer(sd, f"{p}.decoder.layers.{i}", d, enc, cfg.n_heads, cfg.factor, cfg.activation, draw, True, "prob", tops)
    d = _ln(sd, p + ".decoder.norm", d)
    d = _lin(sd, p + ".decoder.projection", d)
    return d[:, -P:, :]


# --------------------------------------------------------------------------------------
# build-defined visual backbone (plain PyTorch twin of the CUDA crop + patch-embed GEMM)
# --------------------------------------------------------------------------------------
def frame_window(H: int, W: int) -> Tuple[float, float, float, float]:
    """(cx, cy, fw, fh) of the "frame" FoV: top-left anchored square of side max(H,W)."""
    side = max(H, W)
    fw, fh = side / W, side / H
    return fw / 2.0, fh / 2.0, fw, fh


def fov_crop(frames: Tensor, centers: Tensor, windows: Tensor, size: int, mean, std) -> Tensor:
    """Bilinear crop/resample + normalise.

    frames [N,3,H,W] (any float dtype, values in [0,1]); centers [N,2] = (cx,cy) and windows
    [N,2] = (fw,fh), all as fractions of the frame.  Output [N,3,size,size] fp32.
    Sampling convention: F.grid_sample(bilinear, zeros, align_corners=False): output pixel
    (i,j) has g = (2j+1)/size - 1; source normalised x = fw*g_x + (2cx-1); source pixel
    x = ((x_n + 1) W - 1)/2.
    """
    N = frames.shape[0]
    theta = torch.zeros(N, 2, 3, dtype=torch.float32, device=frames.device)
    theta[:, 0, 0] = windows[:, 0]
    theta[:, 0, 2] = 2.0 * centers[:, 0] - 1.0
    theta[:, 1, 1] = windows[:, 1]
    theta[:, 1, 2] = 2.0 * centers[:, 1] - 1.0
    grid = F.affine_grid(theta, (N, 3, size, size), align_corners=False)
    out = F.grid_sample(frames.float(), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    m = torch.tensor(mean, dtype=torch.float32, device=frames.device).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32, device=frames.device).view(1, 3, 1, 1)
    return (out - m) / s


def patch_backbone(sd: SD, p: str, frames: Tensor, spec: BackboneSpec,
                   centers: Optional[Tensor] = None) -> Tensor:
    """[N,3,H,W] -> [N,C,G,G] fp32.  centers=None -> "frame" FoV, else gaze-centred window."""
    N, _, H, W = frames.shape
    if centers is None:
        cx, cy, fw, fh = frame_window(H, W)
        centers = torch.tensor([[cx, cy]], dtype=torch.float32).repeat(N, 1)
        windows = torch.tensor([[fw, fh]], dtype=torch.float32).repeat(N, 1)
    else:
        windows = torch.full((N, 2), float(spec.window), dtype=torch.float32)
    img = fov_crop(frames, centers.to(frames.device), windows.to(frames.device), spec.image_size, spec.mean, spec.std)
    return F.conv2d(img, sd[p + ".proj.weight"], sd[p + ".proj.bias"], stride=spec.patch)


# --------------------------------------------------------------------------------------
# Routeformer glue
# --------------------------------------------------------------------------------------
def median_downsample(x: Tensor, target_len: int) -> Tensor:
    """Lower median of consecutive windows (utils/filter.py:5-43). x [B,S,C] -> [B,target_len,C]."""
    B, S, C = x.shape
    if target_len >= S:
        raise ValueError("Target length must be less than the current time steps.")
    stride = S // target_len
    win = x[:, : stride * target_len].reshape(B, target_len, stride, C)
    return win.sort(dim=2).values[:, :, (stride - 1) // 2, :]


def rotate2d(v: Tensor, angle: Tensor) -> Tensor:
    """Rotate [B,L,2] by per-clip angle [B,1,1] (utils/vector.py:6-54)."""
    c, s = torch.cos(angle), torch.sin(angle)
    x, y = v[..., 0:1], v[..., 1:2]
    return torch.cat([c * x - s * y, s * x + c * y], dim=-1)


def motion_features(motion: Tensor, rotate_motion: bool) -> Tuple[Tensor, Tensor]:
    """[B,T,2] -> ([B,T,5], origin_angle [B,1,1])  (routeformer.py:209-225, utils/vector.py:85-111)."""
    angle = torch.atan2(motion[..., 1:2], motion[..., 0:1])
    norm = torch.sqrt(motion[..., 0:1] ** 2 + motion[..., 1:2] ** 2)
    origin = angle[:, -1:, :] if rotate_motion else angle[:, :1, :]
    angle_n = (angle - origin) / math.pi
    accel = F.pad(norm[:, 1:, :] - norm[:, :-1, :], (0, 0, 1, 0))
    if rotate_motion:
        motion = rotate2d(motion, -origin)
    return torch.cat([motion, angle_n, norm, accel], dim=-1), origin


def frame_indices(T: int, rel: int) -> Tensor:
    """Sub-sampled frame indices, never including frame 0 (routeformer.py:415-419)."""
    return torch.flip(torch.arange(T - 1, 0, -rel).long(), dims=[0])


class Routeformer:
    """Functional twin of routeformer/models/routeformer.py:20-533 over a state_dict."""

    def __init__(self, sd: SD, cfg: OracleConfig, spec: Optional[BackboneSpec] = None, fov: str = "frame"):
        self.sd, self.cfg, self.spec, self.fov = sd, cfg, spec, fov
        self.bn_updates: dict = {}
        self.tops: list = []

    # -- visual streams ------------------------------------------------------------
    def _encode_frames(self, frames: Tensor, draw, centers: Optional[Tensor] = None) -> Tensor:
        """routeformer.py:463-491 (non-dropped branch).  [N,3,H,W] -> [N,E]."""
        cfg = self.cfg
        # the backbone plugin returns features in the INPUT dtype (TimmBackbone.py:141-143: fp16 video ->
        # fp16 features), then Routeformer casts to the GPS-backbone dtype (routeformer.py:472-473)
        feats = patch_backbone(self.sd, "video_backbone", frames, self.spec, centers).to(frames.dtype).to(torch.float32)
        tok = feats.permute(0, 2, 3, 1).reshape(feats.shape[0], -1, feats.shape[1])
        tok = torch.cat([tok, -torch.ones_like(tok)[:, :1, :]], dim=1)
        out = perceive_encoder(self.sd, "frame_encoder", tok, 1, cfg, draw, self.tops)
        return out.view(frames.shape[0], cfg.image_embedding_size)

    def _scatter(self, feats: Tensor, idx: Tensor, B: int, T: int) -> Tensor:
        """Zero [B,T,E] with the encoded frames at their time indices (routeformer.py:443-459)."""
        full = torch.zeros(B, T, feats.shape[-1], device=feats.device)
        return full.index_copy(1, idx.to(feats.device), feats.view(B, -1, feats.shape[-1]))

    def _view(self, video: Tensor, rel: int, drop: bool, draw, gaze_xy: Optional[Tensor] = None) -> Tensor:
        B, T = video.shape[:2]
        idx = frame_indices(T, rel)
        if drop:
            feats = torch.zeros(B * len(idx), self.cfg.image_embedding_size, device=video.device)
        else:
            centers = None
            if gaze_xy is not None and self.fov == "gaze":
                centers = gaze_xy[:, idx].reshape(-1, 2).clamp(0.0, 1.0)
            feats = self._encode_frames(video[:, idx].flatten(0, 1), draw, centers)
        return self._scatter(feats, idx, B, T)

    def preprocess(self, batch: dict, training: bool, draw) -> Tuple[Tensor, Optional[Tensor]]:
        """routeformer.py:254-348."""
        cfg, sd = self.cfg, self.sd
        gps = batch["gps"].to(torch.float32)
        motion = gps[:, 1:, :] - gps[:, :-1, :]
        if cfg.normalize_motion:
            motion = (motion - cfg.motion_mean) / cfg.motion_std
        motion = F.pad(motion, (0, 0, 1, 0))
        if not cfg.with_video:
            return motion, None
        streams = []
        if cfg.with_scene:  # routeformer.py:397-461: RIGHT is encoded before LEFT (:427-428)
            drop_l = drop_r = False
            if cfg.view_dropout > 0.0 and training:
                one = bool(torch.rand(1) < cfg.view_dropout)
                drop_l = one and bool(torch.rand(1) < 0.5)
                drop_r = one and not drop_l
            rel = cfg.output_fps // cfg.video_fps
            right = self._view(batch.get("right_video", batch["left_video"]), rel, drop_r or "right_video" not in batch and training, draw)
            left = self._view(batch["left_video"], rel, drop_l, draw)
            streams += [left + sd["left_video_embedding"], right + sd["right_video_embedding"]]
        if cfg.with_gaze:
            drop_g = False
            if cfg.gaze_dropout > 0.0 and training:
                drop_g = bool(torch.rand(1) < cfg.gaze_dropout)
            front = batch["front_video"]
            if drop_g:
                g = torch.zeros(front.shape[0], front.shape[1], cfg.image_embedding_size, device=front.device)
            else:
                gaze = batch["gaze"].to(torch.float32)
                gaze_ds = median_downsample(gaze, cfg.seq_len)
                fv = self._view(front, cfg.output_fps // cfg.gaze_fps, False, draw,
                                gaze_xy=median_downsample(gaze, front.shape[1]) if gaze.shape[1] > front.shape[1] else gaze)
                gq = perceive_encoder(sd, "gaze_encoder", gaze_ds, cfg.seq_len, cfg, draw, self.tops)
                g = perceive_decoder(sd, "gaze_video_decoder", fv, gq, cfg.seq_len, cfg, draw, self.tops)
                g = g[:, : fv.shape[1]]
            streams.append(g + sd["gaze_video_embedding"])
        tokens = torch.cat([*streams, torch.zeros_like(streams[-1]) + sd["video_output_embedding"]], dim=1)
        visual = perceive_encoder(sd, "video_encoder", tokens, cfg.seq_len, cfg, draw, self.tops)
        return motion, visual

    # -- GPS backbone + decode -----------------------------------------------------
    def backbone_forward(self, motion: Tensor, visual: Optional[Tensor], training: bool, draw,
                         pred_len: Optional[int] = None) -> Tensor:
        """routeformer.py:204-252."""
        cfg = self.cfg
        feats, origin = motion_features(motion, cfg.rotate_motion)
        parts = [feats]
        if cfg.with_video:
            parts.append(visual)
        if cfg.only_motion:
            parts[-1] = torch.zeros_like(parts[-1])
        x = torch.cat(parts, dim=-1)
        out = informer(self.sd, "gps_backbone", x, cfg, draw, training, pred_len, self.bn_updates, self.tops)
        if cfg.decoder_mode == "recursive":
            out = out + (x[:, -1:, :] if cfg.dense_prediction else x[:, -1:, :2])
        if cfg.rotate_motion:
            out = torch.cat([rotate2d(out[:, :, :2], origin), out[:, :, 2:]], dim=-1)
        return out

    def postprocess(self, last_gps: Tensor, out: Tensor):
        """routeformer.py:350-395."""
        cfg = self.cfg
        mv = out[:, :, :2]
        if cfg.normalize_motion:
            mv = mv * cfg.motion_std + cfg.motion_mean
        wp = (last_gps + torch.cumsum(mv, dim=1)).to(last_gps.dtype)
        dense = out[:, :, 2 : 2 + cfg.image_embedding_size] if (cfg.with_video and cfg.dense_prediction) else None
        return mv, wp, dense

    def forward(self, batch: dict, training: bool = False, draw=None):
        """routeformer.py:124-202, both branches."""
        draw = draw or CpuRandint()
        cfg = self.cfg
        motion, visual = self.preprocess(batch, training, draw)
        last_gps = batch["gps"][:, -1:, :]
        if training or not cfg.autoregressive:
            out = self.backbone_forward(motion, visual, training, draw)
            _, wp, dense = self.postprocess(last_gps, out)
        else:
            # routeformer.py:164-197: windows of `step` predictions; the window slides over the model's own motion vectors
            # and predicted visual features, the anchor moves to the last predicted position.
            if not cfg.with_video:  # the reference slices its empty visual-feature list here (:187)
                raise TypeError("list indices must be integers or slices, not tuple")
            step, done, wps, denses = cfg.autoregressive_step_size, 0, [], []
            while done < cfg.pred_len:
                out = self.backbone_forward(motion, visual, training, draw, pred_len=step)
                mv, wp_s, dense_s = self.postprocess(last_gps, out)
                wps.append(wp_s)
                denses.append(dense_s)
                motion = torch.cat([motion[:, step:], mv], dim=1)
                last_gps = wp_s[:, -1:, :]
                visual = torch.cat([visual[:, step:], dense_s], dim=1)  # needs dense_prediction, as in the reference
                done += step
            wp = torch.cat(wps, dim=1)[:, : cfg.pred_len]
            dense = torch.cat(denses, dim=1)[:, : cfg.pred_len] if cfg.with_video else None
        return (wp, dense) if cfg.dense_prediction else wp


# --------------------------------------------------------------------------------------
# metrics and loss
# --------------------------------------------------------------------------------------
def ade(pred: Tensor, truth: Tensor) -> Tensor:
    """Mean L2 over every (b,t) (score/error.py:10-29)."""
    return torch.sqrt(((pred - truth) ** 2).sum(-1)).mean()


def fde(pred: Tensor, truth: Tensor) -> Tensor:
    """Frobenius norm of the LAST BATCH ELEMENT's whole error (score/error.py:32-51) -- not a final-step error."""
    return torch.sqrt(((pred[-1] - truth[-1]) ** 2).sum())


def future_discounted_loss(pred: Tensor, truth: Tensor, gamma: float = 0.97, kind: str = "smooth_l1",
                           epsilon: float = 1.0) -> Tensor:
    """losses/future_discounted_mse.py:56-95: mean over all elements of loss(p,g) * gamma^t."""
    T = pred.shape[1]
    w = torch.pow(torch.tensor(gamma), torch.arange(T, device=pred.device)).view(1, T, *([1] * (pred.dim() - 2)))
    err = pred - truth
    if kind == "smooth_l1":
        return (F.smooth_l1_loss(pred, truth, reduction="none") * w).mean()
    err = torch.where(err.abs() < epsilon, torch.zeros_like(err), err)
    return ((err.abs() if kind == "mae" else err ** 2) * w).mean()


# --------------------------------------------------------------------------------------
# the two callers of the path (experiments/full_comparison.py), for one model
# --------------------------------------------------------------------------------------
def training_step(model: "Routeformer", batch: dict, discount: float = 0.97, dense_loss_ratio: float = 0.5,
                  current_epoch: int = 0, draw=None):
    """full_comparison.py:470-532 for a dense-prediction, non-autoregressive model -> (loss, metrics)."""
    draw = draw or CpuRandint()
    inp, target = batch["train"], batch["target"]
    target_gps = target["gps"].to(torch.float32)
    future_gps, future_visual = model.forward(inp, training=True, draw=draw)
    _, target_visual = model.preprocess(target, False, draw)               # :482, eval-mode target pass
    target_visual = target_visual[:, : future_visual.shape[1]].detach()    # :483-485, :497
    trajectory_loss = future_discounted_loss(future_gps, target_gps, discount)
    dense_loss = future_discounted_loss(future_visual, target_visual, discount)
    weight = (dense_loss_ratio * trajectory_loss / max(dense_loss, 1e-6)).detach() if current_epoch >= 10 else 0  # :502-508
    loss = trajectory_loss + weight * dense_loss
    return loss, {"trajectory_loss": trajectory_loss, "dense_loss": dense_loss, "ade": ade(future_gps, target_gps),
                  "fde": fde(future_gps, target_gps), "target_visual": target_visual}


def eval_step(model: "Routeformer", batch: dict, discount: float = 0.97, n_samples: int = 5, draw=None):
    """full_comparison.py:654-679 (the caller seeds torch with 12345 first): mean of five stochastic forwards, per-clip metrics."""
    draw = draw or CpuRandint()
    target_gps = batch["target"]["gps"]
    preds = []
    for _ in range(n_samples):
        out = model.forward(batch["train"], training=False, draw=draw)
        preds.append(out[0] if isinstance(out, tuple) else out)
    mean = torch.stack(preds).mean(dim=0)
    losses = torch.stack([future_discounted_loss(mean[i:i + 1], target_gps[i:i + 1], discount) for i in range(mean.shape[0])])
    ades = torch.stack([ade(mean[i:i + 1], target_gps[i:i + 1]) for i in range(mean.shape[0])])
    fdes = torch.stack([fde(mean[i:i + 1], target_gps[i:i + 1]) for i in range(mean.shape[0])])
    return losses, ades, fdes, mean, torch.stack(preds)


# --------------------------------------------------------------------------------------
# deterministic weights and synthetic batches (shared by golden generation and tests)
# --------------------------------------------------------------------------------------
def fill_state_dict(sd: SD, seed: int) -> SD:
    """Overwrite every float entry of ``sd`` with seed-determined values (key-order independent).

    Matrices ~ N(0, 1/fan_in), LayerNorm/BatchNorm weights ~ 1 + 0.1 N, biases ~ 0.1 N,
    running_var in [0.5, 1.5]; ``pe`` tables and integer buffers are left alone.
    """
    out = {}
    for i, key in enumerate(sorted(sd.keys())):
        t = sd[key]
        if not t.is_floating_point() or key.endswith(".pe"):
            out[key] = t.clone()
            continue
        g = torch.Generator().manual_seed(seed * 100003 + i)
        r = torch.randn(t.shape, generator=g)
        if key.endswith("running_var"):
            v = 0.5 + torch.rand(t.shape, generator=g)
        elif key.endswith("running_mean"):
            v = 0.1 * r
        elif key.endswith("weight") and "norm" in key.split(".")[-2]:
            v = 1.0 + 0.1 * r
        elif key.endswith("_embedding"):
            v = r
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            v = r / math.sqrt(fan_in)
        else:
            v = 0.1 * r
        out[key] = v.to(t.dtype)
    return out


def synthetic_batch(B: int, cfg: OracleConfig, shapes: str = "gem", seed: int = 1, T: Optional[int] = None,
                    scale: float = 1.0, video_dtype=torch.float16) -> dict:
    """GEM- or DR(eye)VE-shaped synthetic clip batch."""
    g = torch.Generator().manual_seed(seed)
    T = T or cfg.seq_len
    mean, std = (1.8332362885457094, 0.9090128501056961) if shapes == "gem" else (4.7068373500451, 2.722694545590219)
    batch = {"gps": torch.cumsum(mean + std * torch.randn(B, T, 2, generator=g), dim=1)}
    if cfg.with_video:
        if shapes == "gem":
            scene, scene_r, front, gaze_n = (86, 384), (86, 384), (324, 326), 40 * T
        elif shapes == "dreyeve":
            scene, scene_r, front, gaze_n = (216, 768), (216, 384), (240, 320), 2 * T
        else:  # "tiny": CPU-cheap shapes for unit tests
            scene, scene_r, front, gaze_n = (20, 48), (20, 48), (36, 34), 4 * T
        sc = lambda hw: (max(4, int(hw[0] * scale)), max(4, int(hw[1] * scale)))
        if cfg.with_scene:
            batch["left_video"] = torch.rand(B, T, 3, *sc(scene), generator=g).to(video_dtype)
            batch["right_video"] = torch.rand(B, T, 3, *sc(scene_r), generator=g).to(video_dtype)
        if cfg.with_gaze:
            batch["front_video"] = torch.rand(B, T, 3, *sc(front), generator=g).to(video_dtype)
            batch["gaze"] = torch.rand(B, gaze_n, 2, generator=g)
    return batch


# --------------------------------------------------------------------------------------
# state_dict layout: key -> tensor template, reference naming
# --------------------------------------------------------------------------------------
def _attn_keys(sd: SD, p: str, D: int) -> None:
    for name in ("query_projection", "key_projection", "value_projection", "out_projection"):
        sd[f"{p}.{name}.weight"] = torch.zeros(D, D)
        sd[f"{p}.{name}.bias"] = torch.zeros(D)


def _ffn_norm_keys(sd: SD, p: str, D: int, dff: int, n_norms: int) -> None:
    sd[f"{p}.conv1.weight"] = torch.zeros(dff, D, 1)
    sd[f"{p}.conv1.bias"] = torch.zeros(dff)
    sd[f"{p}.conv2.weight"] = torch.zeros(D, dff, 1)
    sd[f"{p}.conv2.bias"] = torch.zeros(D)
    for i in range(1, n_norms + 1):
        sd[f"{p}.norm{i}.weight"] = torch.ones(D)
        sd[f"{p}.norm{i}.bias"] = torch.zeros(D)


def _perceive_encoder_keys(sd: SD, p: str, c_in: int, c_out: int, D: int, layers: int, dff: int) -> None:
    sd[f"{p}.value_embedding.tokenConv.weight"] = torch.zeros(D, c_in, 3)
    sd[f"{p}.value_embedding.tokenConv.bias"] = torch.zeros(D)
    sd[f"{p}.position_embedding.pe"] = pe_table(5000, D).unsqueeze(0)
    for i in range(layers):
        _attn_keys(sd, f"{p}.encoder.attn_layers.{i}.attention", D)
        _ffn_norm_keys(sd, f"{p}.encoder.attn_layers.{i}", D, dff, 2)
    sd[f"{p}.encoder.norm.weight"] = torch.ones(D)
    sd[f"{p}.encoder.norm.bias"] = torch.zeros(D)
    sd[f"{p}.projection.weight"] = torch.zeros(c_out, D)
    sd[f"{p}.projection.bias"] = torch.zeros(c_out)


def _perceive_decoder_keys(sd: SD, p: str, c_q: int, D: int, c_out: int, layers: int, dff: int) -> None:
    sd[f"{p}.value_embedding.tokenConv.weight"] = torch.zeros(D, c_q, 3)
    sd[f"{p}.value_embedding.tokenConv.bias"] = torch.zeros(D)
    sd[f"{p}.position_embedding.pe"] = pe_table(5000, D).unsqueeze(0)
    for i in range(layers):
        _attn_keys(sd, f"{p}.decoder.layers.{i}.self_attention", D)
        _attn_keys(sd, f"{p}.decoder.layers.{i}.cross_attention", D)
        _ffn_norm_keys(sd, f"{p}.decoder.layers.{i}", D, dff, 3)
    sd[f"{p}.decoder.norm.weight"] = torch.ones(D)
    sd[f"{p}.decoder.norm.bias"] = torch.zeros(D)
    sd[f"{p}.projection.weight"] = torch.zeros(c_out, D)
    sd[f"{p}.projection.bias"] = torch.zeros(c_out)


def _informer_keys(sd: SD, p: str, cfg: OracleConfig) -> None:
    D = cfg.d_model
    for emb in ("enc_embedding", "dec_embedding"):
        sd[f"{p}.{emb}.value_embedding.tokenConv.weight"] = torch.zeros(D, cfg.enc_in, 3)
        sd[f"{p}.{emb}.position_embedding.pe"] = pe_table(5000, D).unsqueeze(0)
        sd[f"{p}.{emb}.temporal_embedding.embed.weight"] = torch.zeros(D, 1)
    for i in range(cfg.e_layers):
        _attn_keys(sd, f"{p}.encoder.attn_layers.{i}.attention", D)
        _ffn_norm_keys(sd, f"{p}.encoder.attn_layers.{i}", D, cfg.d_ff, 2)
    if cfg.distil:
        for i in range(cfg.e_layers - 1):
            q = f"{p}.encoder.conv_layers.{i}"
            sd[f"{q}.downConv.weight"] = torch.zeros(D, D, 3)
            sd[f"{q}.downConv.bias"] = torch.zeros(D)
            sd[f"{q}.norm.weight"] = torch.ones(D)
            sd[f"{q}.norm.bias"] = torch.zeros(D)
            sd[f"{q}.norm.running_mean"] = torch.zeros(D)
            sd[f"{q}.norm.running_var"] = torch.ones(D)
            sd[f"{q}.norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd[f"{p}.encoder.norm.weight"] = torch.ones(D)
    sd[f"{p}.encoder.norm.bias"] = torch.zeros(D)
    for i in range(cfg.d_layers):
        _attn_keys(sd, f"{p}.decoder.layers.{i}.self_attention", D)
        _attn_keys(sd, f"{p}.decoder.layers.{i}.cross_attention", D)
        _ffn_norm_keys(sd, f"{p}.decoder.layers.{i}", D, cfg.d_ff, 3)
    sd[f"{p}.decoder.norm.weight"] = torch.ones(D)
    sd[f"{p}.decoder.norm.bias"] = torch.zeros(D)
    sd[f"{p}.decoder.projection.weight"] = torch.zeros(cfg.c_out, D)
    sd[f"{p}.decoder.projection.bias"] = torch.zeros(cfg.c_out)


def state_dict_template(cfg: OracleConfig, spec: Optional[BackboneSpec] = None) -> SD:
    """Reference-layout state_dict (keys, shapes, dtypes) of Routeformer for ``cfg`` (routeformer.py:57-117)."""
    sd: SD = {}
    E, H = cfg.image_embedding_size, cfg.encoder_hidden_size
    if cfg.with_video:
        assert spec is not None
        sd["video_backbone.proj.weight"] = torch.zeros(spec.channels, 3, spec.patch, spec.patch)
        sd["video_backbone.proj.bias"] = torch.zeros(spec.channels)
        _perceive_encoder_keys(sd, "frame_encoder", spec.channels, E, cfg.perceive_d_model, cfg.encoder_layers, cfg.encoder_d_ff)
        for name in ("left_video_embedding", "right_video_embedding", "gaze_video_embedding", "video_output_embedding"):
            sd[name] = torch.zeros(1, 1, E)
        _perceive_encoder_keys(sd, "video_encoder", E, H, cfg.perceive_d_model, cfg.encoder_layers, cfg.encoder_d_ff)
        if cfg.with_gaze:
            _perceive_encoder_keys(sd, "gaze_encoder", 2, H, cfg.perceive_d_model, cfg.encoder_layers, cfg.encoder_d_ff)
            _perceive_decoder_keys(sd, "gaze_video_decoder", H, H, H, cfg.cross_modal_decoder_layers, cfg.encoder_d_ff)
    _informer_keys(sd, "gps_backbone", cfg)
    return sd
