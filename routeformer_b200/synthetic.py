"""Synthetic GEM- / DR(eye)VE-shaped clip batches and the paper configuration (benchmarks, examples).

Shapes follow the reference's data pipeline:
  GEM      (experiments/full_comparison.py:113-128, routeformer/io/dataset.py:1324-1338,1439-1523): scene views 86x384,
           front (eye-tracker world camera) 324x326, gaze 200 Hz x 8 s;
  DR(eye)VE (full_comparison.py:107-110,534-551, io/dataset_dreyeve.py:1219-1227): roof camera 216x768 / 216x384,
           eye-tracker camera 240x320, two gaze samples per frame.
Values: gps = cumulative sum of N(mean, std) per-step displacements (the datasets' motion statistics), frames U(0,1) in fp16
(io/dataset.py:1522 divides by 255 on the host), gaze U(0,1) normalised image coordinates.
"""
from __future__ import annotations

from typing import Optional

import torch

from .backbone import PatchEmbedBackbone
from .config import GPSBackboneConfig, PatchBackboneConfig, RouteformerConfig
from .informer import Informer

SHAPES = {
    # name: (left/scene (H, W), right (H, W), front (H, W), gaze samples per output step, motion mean, motion std)
    "gem": ((86, 384), (86, 384), (324, 326), 40, 1.8332362885457094, 0.9090128501056961),
    "dreyeve": ((216, 768), (216, 384), (240, 320), 2, 4.7068373500451, 2.722694545590219),
    "tiny": ((20, 48), (20, 48), (36, 34), 4, 4.7068373500451, 2.722694545590219),
}


def synthetic_batch(B: int, T: int = 40, shapes: str = "gem", seed: int = 1, with_scene: bool = True, with_gaze: bool = True,
                    scale: float = 1.0, video_dtype=torch.float16) -> dict:
    """Host batch in the reference layout: gps [B,T,2], *_video [B,T,3,H,W], gaze [B,n,2]."""
    g = torch.Generator().manual_seed(seed)
    scene, scene_r, front, gaze_per_step, mean, std = SHAPES[shapes]
    batch = {"gps": torch.cumsum(mean + std * torch.randn(B, T, 2, generator=g), dim=1)}
    sc = lambda hw: (max(4, int(hw[0] * scale)), max(4, int(hw[1] * scale)))

    def frames(hw):
        # generated clip by clip: a 64-clip GEM batch is 2.6 GB in fp16 (twice that as the fp32 draws it is cast from)
        out = torch.empty(B, T, 3, *sc(hw), dtype=video_dtype)
        for b in range(B):
            out[b] = torch.rand(T, 3, *sc(hw), generator=g).to(video_dtype)
        return out

    if with_scene:
        batch["left_video"] = frames(scene)
        batch["right_video"] = frames(scene_r)
    if with_gaze:
        batch["front_video"] = frames(front)
        batch["gaze"] = torch.rand(B, gaze_per_step * T, 2, generator=g)
    return batch


def synthetic_targets(batch: dict, pred_len: int = 30, embedding: int = 64, shapes: str = "gem", seed: int = 1001):
    """(future waypoints [B,P,2] continuing the track, dense feature targets [B,P,E])."""
    g = torch.Generator().manual_seed(seed)
    _, _, _, _, mean, std = SHAPES[shapes]
    B = batch["gps"].shape[0]
    return (batch["gps"][:, -1:, :] + torch.cumsum(mean + std * torch.randn(B, pred_len, 2, generator=g), 1),
            torch.randn(B, pred_len, embedding, generator=g))


def paper_model(fov: str = "gaze", rotate_motion: bool = False, view_dropout: float = 0.0, gaze_dropout: float = 0.0,
                feature_dropout: float = 0.0, image_size: int = 256, patch: int = 32, channels: int = 1024, window: float = 0.5,
                seq_len: int = 40, pred_len: int = 30):
    """Routeformer at the paper configuration (ROUTEFORMER_CONFIG_SWINV2_GAZE, experiments/full_comparison.py:159-296) with the
    build-defined random-init patch backbone in place of the un-vendored timm SwinV2-B (same 8x8x1024 feature map)."""
    from .routeformer import Routeformer

    g = GPSBackboneConfig(seq_len=seq_len, label_len=seq_len, pred_len=pred_len, factor=4, distil=True, dropout=0.0, activation="relu",
                          d_model=832, n_heads=8, e_layers=6, d_layers=1, d_ff=3328)
    vb = PatchBackboneConfig(image_size=image_size, patch=patch, channels=channels, fov=fov, window=window)
    rc = RouteformerConfig(gps_backbone_config=g, video_backbone_config=vb, decoder_mode="smart", with_video=True, with_gaze=True,
                           dense_prediction=True, dense_loss_ratio=0.5, image_embedding_size=64, encoder_hidden_size=64,
                           encoder_heads=8, encoder_layers=8, encoder_d_ff=256, cross_modal_decoder_heads=8,
                           cross_modal_decoder_layers=2, rotate_motion=rotate_motion, view_dropout=view_dropout,
                           gaze_dropout=gaze_dropout, feature_dropout=feature_dropout)
    return Routeformer(rc, gps_backbone=Informer, video_backbone=PatchEmbedBackbone)
