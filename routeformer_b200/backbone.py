"""Visual backbone plugin interface + the build-defined random-init encoder.

`VideoBackboneModule` keeps the reference's plugin contract (routeformer/models/video_backbone/config.py:45-52):
constructed as `cls(configs=VideoBackboneConfig)`, exposes `output_feature_shape -> (C, Hf, Wf)` and
`forward([N,3,H,W]) -> [N,C,Hf,Wf]` in the input dtype.

`PatchEmbedBackbone` replaces the un-vendored timm SwinV2 (TimmBackbone.py:13-184; "parity unpinned", random-init
encoders per BASELINE.json): gaze-centred / whole-frame FoV crop kernel -> patch-embedding GEMM on tcgen05.
Its plain-PyTorch twin (the oracle) is oracle.routeformer_oracle.patch_backbone.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .config import PatchBackboneConfig


class VideoBackboneModule(ABC, nn.Module):
    current_epoch = 0

    @property
    @abstractmethod
    def output_feature_shape(self) -> tuple:
        """(C, Hf, Wf) of the feature map."""

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


def frame_window(H: int, W: int) -> Tuple[float, float, float, float]:
    """(cx, cy, fw, fh) of the reference-like FoV: zero-pad bottom/right to a square, then resize (TimmBackbone.py:164-176)."""
    side = max(H, W)
    fw, fh = side / W, side / H
    return fw / 2.0, fh / 2.0, fw, fh


class PatchEmbedBackbone(VideoBackboneModule):
    def __init__(self, configs: Optional[PatchBackboneConfig] = None):
        super().__init__()
        self.configs = configs if isinstance(configs, PatchBackboneConfig) else PatchBackboneConfig()
        c = self.configs
        assert c.image_size % c.patch == 0 and c.patch % 4 == 0
        self.proj = nn.Conv2d(3, c.channels, c.patch, c.patch)
        self.train_backbone = bool(c.train_backbone)
        self.proj.requires_grad_(self.train_backbone)
        self._cache = {}

    @property
    def output_feature_shape(self) -> tuple:
        return (self.configs.channels, self.configs.grid, self.configs.grid)

    @property
    def tokens_per_frame(self) -> int:
        return self.configs.grid ** 2 + 1

    # -- helpers ---------------------------------------------------------------------------------
    def _frame_ids(self, B: int, T: int, t_idx: torch.Tensor, device) -> torch.Tensor:
        key = ("ids", B, T, tuple(t_idx.tolist()), str(device))
        if key not in self._cache:
            ids = (torch.arange(B).view(B, 1) * T + t_idx.view(1, -1)).reshape(-1).to(torch.int32)
            self._cache[key] = ids.to(device)
        return self._cache[key]

    def _frame_fov(self, n: int, H: int, W: int, device):
        key = ("fov", n, H, W, str(device))
        if key not in self._cache:
            cx, cy, fw, fh = frame_window(H, W)
            self._cache[key] = (torch.tensor([[cx, cy]], dtype=torch.float32).repeat(n, 1).to(device),
                                torch.tensor([[fw, fh]], dtype=torch.float32).repeat(n, 1).to(device))
        return self._cache[key]

    def encode_views(self, views: List[dict]) -> torch.Tensor:
        """views: [{video [B,T,3,H,W] (contiguous, fp16/fp32/u8), t_idx LongTensor[F] (cpu), centers [B*F,2] or None}]
        -> token buffer [sum_v B*F*(G*G+1), C] fp32: per frame G*G patch features then the constant -1 token
        (routeformer.py:478-487), ready for the frame encoder."""
        c = self.configs
        G, S, p, C = c.grid, c.image_size, c.patch, c.channels
        dev = views[0]["video"].device
        n_total = sum(v["video"].shape[0] * len(v["t_idx"]) for v in views)
        # fp16 clips: crop to fp16 patches and multiply in fp16 with fp32 accumulation, as the reference backbone does under
        # autocast (TimmBackbone.py:106-145) -- half the patch bytes and twice the MMA rate of the TF32 path; other dtypes: TF32.
        # uint8 clips (raw frames staged on the device: half the host->device bytes) are converted in the crop kernel exactly as the
        # reference's loader converts them on the host (fp16(v / 255), io/dataset.py:1505-1522) and then follow the fp16 path.
        half = all(v["video"].dtype in (torch.float16, torch.uint8) for v in views) and not self.proj.weight.requires_grad
        # bf16 operand mode (ops.bf16_operands): bf16 patches x bf16 weight; in inference the features are also STORED in bf16 for
        # the frame encoder's token convolution (the autograd path keeps fp32 features for the weight gradient)
        bf16 = half and ops.BF16_MODE
        op_dtype = torch.bfloat16 if bf16 else (torch.float16 if half else torch.float32)
        patches = torch.empty(n_total * G * G, 3 * p * p, device=dev, dtype=op_dtype)
        row = 0
        round_f16 = True
        for v in views:
            video = v["video"]
            assert video.is_contiguous(), "video tensors must be contiguous [B,T,3,H,W]"
            B, T, _, H, W = video.shape
            n = B * len(v["t_idx"])
            ids = self._frame_ids(B, T, v["t_idx"], dev)
            if v.get("centers") is not None and c.fov == "gaze":
                centers = v["centers"].to(torch.float32).contiguous()
                windows = torch.full((n, 2), float(c.window), device=dev, dtype=torch.float32)
            else:
                centers, windows = self._frame_fov(n, H, W, dev)
            ops.fov_crop(video, centers, windows, S, c.mean, c.std, patch=p, frame_ids=ids, n_frames=n,
                         out=patches[row * G * G:(row + n) * G * G], out_dtype=op_dtype, u8_as_f16=half)
            round_f16 = round_f16 and video.dtype in (torch.float16, torch.uint8)
            row += n
        # the plugin returns features in the input dtype (fp16 video -> fp16 features, TimmBackbone.py:141-143)
        if torch.is_grad_enabled() and self.proj.weight.requires_grad:  # train_backbone: fp32 patches, TF32 GEMM, weight gradient
            from . import functional as Fn
            return Fn.PatchEmbed.apply(patches, self.proj.weight, self.proj.bias, G * G, round_f16)
        bf16_features = bf16 and not torch.is_grad_enabled()
        tokens = torch.empty(n_total * (G * G + 1), C, device=dev, dtype=torch.bfloat16 if bf16_features else torch.float32)
        tokens.view(n_total, G * G + 1, C)[:, G * G, :] = -1.0
        ops.gemm(patches, self._weight_matrix(op_dtype), tokens, bias=self.proj.bias, out_group=(G * G, G * G + 1, 0),
                 round_f16=round_f16 and not bf16)
        return tokens

    def _weight_matrix(self, dtype: torch.dtype) -> torch.Tensor:
        """[C, 3*p*p] view of the projection weight; for the 16-bit paths a cached copy, refreshed when the weight changes."""
        w = self.proj.weight
        mat = w.view(w.shape[0], -1)
        if dtype == torch.float32:
            return mat
        key = (w.data_ptr(), w._version, dtype)
        cached = getattr(self, "_w16", None)
        if cached is None or cached[0] != key:
            self._w16 = (key, mat.detach().to(dtype).contiguous())
        return self._w16[1]

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """Plugin API: [N,3,H,W] -> [N,C,G,G] in the input dtype, whole-frame FoV."""
        c = self.configs
        N = images.shape[0]
        video = images.contiguous().unsqueeze(0)
        tokens = self.encode_views([{"video": video, "t_idx": torch.arange(N), "centers": None}])
        G, C = c.grid, c.channels
        feats = tokens.view(N, G * G + 1, C)[:, :G * G, :].reshape(N, G, G, C).permute(0, 3, 1, 2)
        return feats.to(images.dtype) if images.dtype.is_floating_point else feats
