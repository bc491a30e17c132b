"""Informer GPS backbone ("trajectory decoder head") on the CUDA library.

Drop-in for routeformer/models/gps_backbone/Informer.py:18-167: same constructor (`configs: GPSBackboneConfig`),
mutable `pred_len`, `forward(x [B,L,enc_in]) -> [B,pred_len,c_out]`, same sub-module names / state_dict layout
(gps_backbone/layers/{Embedding,SelfAttentionFamily,TransformerEncoderDecoder}.py).
Layout quirk reproduced: the attention context stays [B,H,L,dh] and is re-read as [B,L,H*dh]
(SelfAttentionFamily.py:165,192) -- the kernel writes that layout directly (RF_LAYOUT_BHLD).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as tF

from . import functional as Fn
from .layers import (AttentionLayer, Decoder, DecoderLayer, Encoder, EncoderLayer, LiveIndexSource, PositionalEmbedding,
                     TokenEmbedding, _site)


class TimeFeatureEmbedding(nn.Module):
    def __init__(self, d_model, freq="m"):
        super().__init__()
        d_inp = {"h": 4, "t": 5, "s": 6, "m": 1, "a": 1, "w": 2, "d": 3, "b": 3}[freq]
        if d_inp != 1:
            raise NotImplementedError("only the scalar time feature (freq 'm'/'a') of the Routeformer path is implemented")
        self.embed = nn.Linear(d_inp, d_model, bias=False)


class DataEmbedding(nn.Module):
    """value (circular conv, no bias) + time feature (t * w) + positional table (Embedding.py:111-126)."""

    def __init__(self, c_in, d_model, embed_type="timeF", freq="m", dropout=0.0):
        super().__init__()
        if embed_type != "timeF":
            raise NotImplementedError("only embed='timeF' (the Routeformer configuration) is implemented")
        self.value_embedding = TokenEmbedding(c_in, d_model, bias=False)
        self.position_embedding = PositionalEmbedding(d_model)
        self.temporal_embedding = TimeFeatureEmbedding(d_model, freq)
        self.p_drop = dropout

    def embed(self, x2, n, L, name=""):
        h = Fn.CircularConv3.apply(x2, self.value_embedding.tokenConv.weight, None, self.position_embedding.table(),
                                   self.temporal_embedding.embed.weight, n, L, 1)
        if self.training and self.p_drop > 0.0:  # Embedding.py:126
            h = Fn.Dropout.apply(h, self.p_drop, _site(h.device, self.p_drop, name + ".dropout", h.shape[0], h.shape[1]))
        return h


class ConvLayer(nn.Module):
    """Distilling block: circular Conv1d(k=3, pad=2) -> BatchNorm1d -> ELU -> MaxPool1d(3,2,1) (TransformerEncoderDecoder.py:9-29)."""

    def __init__(self, c_in):
        super().__init__()
        self.downConv = nn.Conv1d(c_in, c_in, kernel_size=3, padding=2, padding_mode="circular")
        self.norm = nn.BatchNorm1d(c_in)

    def run(self, h2, B, L):
        z = Fn.CircularConv3.apply(h2, self.downConv.weight, self.downConv.bias, None, None, B, L, 2)
        bn = self.norm
        out = Fn.DistilTail.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, B, L + 2, self.training)
        if self.training:
            bn.num_batches_tracked += 1
        return out, (L + 1) // 2 + 1


class Informer(nn.Module):
    def __init__(self, configs):
        super().__init__()
        if configs.output_attention:
            raise NotImplementedError("output_attention=True is not supported by the fused attention kernel")
        self.pred_len = configs.pred_len
        self.output_attention = configs.output_attention
        self.smart_decoder = configs.smart_decoder
        self.enc_in = configs.enc_in
        self.c_out = configs.c_out
        D, H, f = configs.d_model, configs.n_heads, configs.factor
        self.enc_embedding = DataEmbedding(configs.enc_in, D, configs.embed, configs.freq, configs.dropout)
        self.dec_embedding = DataEmbedding(configs.dec_in, D, configs.embed, configs.freq, configs.dropout)
        self.encoder = Encoder(
            [EncoderLayer(AttentionLayer("prob", f, D, H, informer_layout=True), D, configs.d_ff, dropout=configs.dropout,
                          activation=configs.activation) for _ in range(configs.e_layers)],
            [ConvLayer(D) for _ in range(configs.e_layers - 1)] if configs.distil else None,
            norm_layer=nn.LayerNorm(D))
        self.decoder = Decoder(
            [DecoderLayer(AttentionLayer("prob_masked", f, D, H, informer_layout=True),
                          AttentionLayer("prob", f, D, H, informer_layout=True), D, configs.d_ff, dropout=configs.dropout,
                          activation=configs.activation) for _ in range(configs.d_layers)],
            norm_layer=nn.LayerNorm(D), projection=nn.Linear(D, configs.c_out, bias=True))

    @property
    def padded_in(self) -> int:
        return (self.enc_in + 3) // 4 * 4

    def run(self, x_pad: torch.Tensor, draw, record=None, name="gps_backbone") -> torch.Tensor:
        """x_pad [B,T,padded_in] (columns >= enc_in zero) -> [B, pred_len, c_out]."""
        B, T, ld = x_pad.shape
        P = self.pred_len
        x_dec = Fn.DecoderInput.apply(x_pad, P, self.smart_decoder)
        h = self.enc_embedding.embed(x_pad.view(B * T, ld), B, T, f"{name}.enc_embedding")
        L = T
        n_layers = len(self.encoder.attn_layers)
        for i, layer in enumerate(self.encoder.attn_layers):
            h = layer.run(h, B, L, draw, record, f"{name}.encoder.attn_layers.{i}")
            if self.encoder.conv_layers is not None and i < n_layers - 1:
                h, L = self.encoder.conv_layers[i].run(h, B, L)
        enc = Fn.LayerNorm.apply(h, self.encoder.norm.weight, self.encoder.norm.bias)
        d = self.dec_embedding.embed(x_dec.view(B * (T + P), ld), B, T + P, f"{name}.dec_embedding")
        for i, layer in enumerate(self.decoder.layers):
            d = layer.run(d, enc, B, T + P, L, draw, record, f"{name}.decoder.layers.{i}")
        d = d.view(B, T + P, -1)[:, T:, :].reshape(B * P, -1)  # only the last P rows are projected / returned (Informer.py:164-167)
        d = Fn.LayerNorm.apply(d, self.decoder.norm.weight, self.decoder.norm.bias)
        out = Fn.Linear.apply(d, self.decoder.projection.weight, self.decoder.projection.bias)
        return out.view(B, P, self.c_out)

    def forward(self, x: torch.Tensor, draw=None) -> torch.Tensor:
        draw = draw or LiveIndexSource(x.device)
        x = x.to(torch.float32)
        if x.shape[-1] != self.padded_in:
            x = tF.pad(x, (0, self.padded_in - x.shape[-1]))
        return self.run(x.contiguous(), draw)
