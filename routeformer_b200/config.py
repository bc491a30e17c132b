"""Configuration objects of the drop-in API.

Same names, fields, defaults and derived properties as the reference so `experiments/full_comparison.py`
can build them unchanged:
  BaseConfig            <- routeformer/utils/config.py:7-33
  GPSBackboneConfig     <- routeformer/models/gps_backbone/config.py:9-75
  VideoBackboneConfig   <- routeformer/models/video_backbone/config.py:10-27
  RouteformerConfig     <- routeformer/models/config.py:10-107
`PatchBackboneConfig` is new: it configures the build-defined random-init visual encoder (FoV crop + patch
embedding) that stands in for the un-vendored timm model.
"""
from __future__ import annotations

import copy
from argparse import Namespace
from dataclasses import dataclass, field
from typing import Literal, Tuple


class BaseConfig(Namespace):
    """argparse.Namespace with item access, deep `copy()` and `override(**kw)` that re-runs `__post_init__`."""

    def __getitem__(self, item):
        return getattr(self, item)

    def get(self, item, default):
        return getattr(self, item, default)

    def __copy__(self):
        return copy.deepcopy(self)

    def copy(self):
        return copy.deepcopy(self)

    def override(self, **kwargs):
        new = self.copy()
        for key, value in kwargs.items():
            setattr(new, key, value)
        if hasattr(new, "__post_init__"):
            new.__post_init__()
        return new


@dataclass
class GPSBackboneConfig(BaseConfig):
    seq_len: int
    label_len: int
    pred_len: int
    embed: str = "timeF"
    freq: str = "m"
    d_model: int = 128
    n_heads: int = 8
    e_layers: int = 2
    d_layers: int = 1
    d_ff: int = 512
    moving_avg: int = 25
    factor: int = 1
    distil: bool = True
    dropout: float = 0.1
    activation: str = "gelu"
    individual: bool = False
    # filled in by RouteformerConfig.__post_init__
    output_attention: bool = field(init=False)
    with_video: bool = field(init=False)
    with_gaze: bool = field(init=False)
    dense_prediction: bool = field(init=False)
    encoder_hidden_size: int = field(init=False)
    image_embedding_size: int = field(init=False)
    output_fps: int = field(init=False)
    dense_loss_ratio: float = field(init=False)
    discount_factor: dict = field(init=False)
    smart_decoder: bool = field(init=False)
    _enc_in: int = None
    _c_out: int = None

    @property
    def enc_in(self) -> int:
        """5 motion channels (dx, dy, angle, speed, acceleration) + the fused visual embedding if any."""
        if self._enc_in is not None:
            return self._enc_in
        return 5 + (self.encoder_hidden_size if self.with_video else 0)

    @property
    def c_out(self) -> int:
        if self._c_out is not None:
            return self._c_out
        return self.enc_in - 3 if self.dense_prediction else 2

    @property
    def dec_in(self) -> int:
        return self.enc_in


@dataclass
class VideoBackboneConfig(BaseConfig):
    cache_dir: str = None
    train_backbone: bool = False
    backbone_minibatch_size: int = 4
    torchcache_enabled: bool = True
    torchcache_persistent_module_hash: str = None
    torchcache_max_persistent_cache_size: int = 200e9
    torchcache_max_memory_cache_size: int = 20e9

    def __post_init__(self):
        if self.torchcache_enabled and self.train_backbone:
            raise ValueError("torchcache_enabled and train_backbone cannot both be True.")


@dataclass
class PatchBackboneConfig(VideoBackboneConfig):
    """Build-defined encoder: FoV crop/resample -> normalise -> Conv2d(3, channels, patch, stride=patch)."""

    torchcache_enabled: bool = False
    image_size: int = 256
    patch: int = 32
    channels: int = 1024
    fov: Literal["frame", "gaze"] = "frame"  # "frame": pad-to-square + resize (reference-like); "gaze": window on the fixation
    window: float = 0.5                      # window size (fraction of the frame) in "gaze" mode
    mean: Tuple[float, float, float] = (0.485, 0.456, 0.406)
    std: Tuple[float, float, float] = (0.229, 0.224, 0.225)

    @property
    def grid(self) -> int:
        return self.image_size // self.patch


@dataclass
class RouteformerConfig(BaseConfig):
    gps_backbone_config: GPSBackboneConfig
    video_backbone_config: VideoBackboneConfig = None
    output_attention: bool = False
    with_video: bool = None
    with_gaze: bool = False
    with_scene: bool = True
    discount_factor: dict = field(default_factory=lambda: {0: 0.9})
    decoder_mode: Literal["vanilla", "recursive", "smart"] = "vanilla"
    rotate_motion: bool = False
    loss_function: Literal["mse", "mae", "smooth_l1"] = "smooth_l1"
    epsilon: float = None
    visual_epsilon: float = None
    autoregressive: bool = False
    autoregressive_step_size: int = 1
    dense_prediction: bool = False
    dense_loss_ratio: float = 0.25
    video_fps: int = 1
    gaze_fps: int = 1
    encoder_hidden_size: int = 64
    encoder_heads: int = 8
    encoder_layers: int = 2
    encoder_d_ff: int = 64
    cross_modal_decoder_heads: int = 8
    cross_modal_decoder_layers: int = 1
    normalize_motion: bool = False
    motion_mean: float = 0.0
    motion_std: float = 1.0
    motion_noise: float = 0.0
    view_dropout: float = 0.0
    gaze_dropout: float = 0.0
    feature_dropout: float = 0.0
    image_embedding_size: int = 128
    # training knobs carried for the experiment driver; unused by the model
    lr: float = 5e-4
    wd: float = 0
    optimizer: str = "Adam"
    batch_size: int = 32
    min_pci: float = 0.0
    step_size: int = 1
    epochs: int = 100
    output_fps: int = 5
    gopro_scaling_factor: float = 1.0
    front_scaling_factor: float = 1.0
    num_workers: int = 0
    use_cache: bool = False
    cache_dir: str = None
    _only_motion: bool = False

    def __post_init__(self, **kwargs):
        assert self.output_fps % self.video_fps == 0, "Video FPS must be a divisor of the output FPS"
        assert self.output_fps % self.gaze_fps == 0, "Gaze FPS must be a divisor of the output FPS"
        if self.with_video is None:
            self.with_video = self.video_backbone_config is not None
        if self.with_gaze:
            assert self.with_video, "Gaze backbone requires video backbone to be used"
        child = self.gps_backbone_config
        child.output_attention = self.output_attention
        child.with_video = self.with_video
        child.with_gaze = self.with_gaze
        child.dense_prediction = self.dense_prediction
        child.image_embedding_size = self.image_embedding_size
        child.encoder_hidden_size = self.encoder_hidden_size
        child.output_fps = self.output_fps
        child.dense_loss_ratio = self.dense_loss_ratio
        child.discount_factor = self.discount_factor
        child.smart_decoder = self.decoder_mode == "smart"
