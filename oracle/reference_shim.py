"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference modules from /root/reference.

The reference cannot be imported as a package here (lightning, timm, torchcache, av, ... are
not installed), so the few third-party names the hot path touches are shimmed in
``sys.modules`` and the reference packages are registered without running their
``__init__.py`` (recipe: SURVEY.md Appendix A).  Only used in THIS container (the reference is
not on the GPU box): to validate oracle/routeformer_oracle.py and to generate tests/golden/.

Never call ``torch.set_float32_matmul_precision("medium")`` in a process that uses this.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("ROUTEFORMER_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "routeformer", "models", "routeformer.py"))


_loaded = None


def load():
    """Returns a namespace with the reference classes/functions of the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not mounted at {REF}")

    lightning = types.ModuleType("lightning")

    class LightningModule(nn.Module):
        current_epoch = 0

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

    lightning.LightningModule = LightningModule
    sys.modules["lightning"] = lightning
    tc = types.ModuleType("torchcache")
    tc.torchcache = lambda *a, **k: (lambda cls: cls)
    sys.modules["torchcache"] = tc

    def pkg(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m

    pkg("routeformer", f"{REF}/routeformer")
    pkg("routeformer.io")
    ds = types.ModuleType("routeformer.io.dataset")
    ds.Data = ds.Item = dict
    sys.modules["routeformer.io.dataset"] = ds
    pkg("routeformer.utils", f"{REF}/routeformer/utils")
    pkg("routeformer.models", f"{REF}/routeformer/models")
    vb = pkg("routeformer.models.video_backbone", f"{REF}/routeformer/models/video_backbone")
    from routeformer.models.video_backbone.config import VideoBackboneConfig, VideoBackboneModule

    vb.VideoBackboneConfig, vb.VideoBackboneModule = VideoBackboneConfig, VideoBackboneModule
    from routeformer.models.config import RouteformerConfig
    from routeformer.models.cross_modal_transformer import PerceiveDecoder, PerceiveEncoder
    from routeformer.models.gps_backbone import GPSBackboneConfig, Informer
    from routeformer.models.routeformer import Routeformer
    from routeformer.utils.filter import median_downsampler

    def by_path(path, name):
        sp = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(sp)
        sp.loader.exec_module(m)
        return m

    err = by_path(f"{REF}/routeformer/score/error.py", "ref_error")
    loss = by_path(f"{REF}/routeformer/losses/future_discounted_mse.py", "ref_loss")

    ns = types.SimpleNamespace(
        Routeformer=Routeformer, RouteformerConfig=RouteformerConfig, GPSBackboneConfig=GPSBackboneConfig,
        Informer=Informer, PerceiveEncoder=PerceiveEncoder, PerceiveDecoder=PerceiveDecoder,
        VideoBackboneConfig=VideoBackboneConfig, VideoBackboneModule=VideoBackboneModule,
        median_downsampler=median_downsampler, ade=err.ade, fde=err.fde,
        FutureDiscountedLoss=loss.FutureDiscountedLoss,
    )
    _loaded = ns
    return ns


def make_ref_backbone_class(spec):
    """nn.Module twin of the build-defined patch-embed encoder, as a reference VideoBackboneModule."""
    ref = load()
    from oracle import routeformer_oracle as O

    class RefPatchBackbone(ref.VideoBackboneModule):
        def __init__(self, configs=None):
            super().__init__()
            self.proj = nn.Conv2d(3, spec.channels, spec.patch, spec.patch)

        @property
        def output_feature_shape(self):
            return (spec.channels, spec.grid, spec.grid)

        def forward(self, images):
            sd = {"b.proj.weight": self.proj.weight, "b.proj.bias": self.proj.bias}
            return O.patch_backbone(sd, "b", images, spec).to(images.dtype)

    return RefPatchBackbone


def build_reference_model(cfg, spec=None):
    """Reference Routeformer for an OracleConfig (mirrors experiments/full_comparison.py:159-296)."""
    ref = load()
    gps_cfg = ref.GPSBackboneConfig(
        seq_len=cfg.seq_len, label_len=cfg.seq_len, pred_len=cfg.pred_len, embed="timeF", freq="m",
        moving_avg=25, factor=cfg.factor, distil=cfg.distil, dropout=0.0, activation=cfg.activation,
        individual=False, d_model=cfg.d_model, n_heads=cfg.n_heads, e_layers=cfg.e_layers,
        d_layers=cfg.d_layers, d_ff=cfg.d_ff,
    )
    rcfg = ref.RouteformerConfig(
        gps_backbone_config=gps_cfg, decoder_mode=cfg.decoder_mode, rotate_motion=cfg.rotate_motion,
        normalize_motion=cfg.normalize_motion, motion_mean=cfg.motion_mean, motion_std=cfg.motion_std,
        with_video=cfg.with_video, with_scene=cfg.with_scene, with_gaze=cfg.with_gaze,
        video_fps=cfg.video_fps, gaze_fps=cfg.gaze_fps, dense_prediction=cfg.dense_prediction,
        image_embedding_size=cfg.image_embedding_size, encoder_hidden_size=cfg.encoder_hidden_size,
        encoder_heads=cfg.encoder_heads, encoder_layers=cfg.encoder_layers, encoder_d_ff=cfg.encoder_d_ff,
        cross_modal_decoder_heads=cfg.cross_modal_decoder_heads,
        cross_modal_decoder_layers=cfg.cross_modal_decoder_layers,
        view_dropout=cfg.view_dropout, gaze_dropout=cfg.gaze_dropout, feature_dropout=0.0,
        autoregressive=cfg.autoregressive, autoregressive_step_size=cfg.autoregressive_step_size,
        output_fps=cfg.output_fps, video_backbone_config=ref.VideoBackboneConfig() if cfg.with_video else None,
    )
    backbone = make_ref_backbone_class(spec) if cfg.with_video else None
    return ref.Routeformer(rcfg, gps_backbone=ref.Informer, video_backbone=backbone)
