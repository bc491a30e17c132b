"""Routeformer on the CUDA library -- drop-in for routeformer/models/routeformer.py:20-533.

Same constructor (`Routeformer(configs, gps_backbone=Informer, video_backbone=None)`), same public methods
(`forward`, `preprocess_batch`, `postprocess_batch`), same sub-module names and state_dict layout, same CPU-RNG
consumption (ProbSparse key sampling with torch.randint, view/gaze dropout with torch.rand) in the same order.

What is different underneath:
  * all CPU random draws of a forward are made up-front, in reference order, and uploaded with ONE copy;
  * the three camera views (right, left, front) share the frame encoder, so their frames are encoded in ONE
    batched pass (per-view sampled-key tables are kept apart through the kernel's `idx_group`);
  * the visual backbone runs as FoV-crop kernel -> tcgen05 patch-embedding GEMM writing the token buffer of
    the frame encoder directly (no permute / cat / scatter copies).
"""
from __future__ import annotations

from typing import List, Optional, Type

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .arena import Arena
from .backbone import PatchEmbedBackbone, VideoBackboneModule
from .config import RouteformerConfig
from .informer import Informer
from .layers import LiveIndexSource, PerceiveDecoder, PerceiveEncoder, PlannedIndexSource, sparse_budget


def frame_indices(T: int, rel: int) -> torch.Tensor:
    """Sub-sampled frame times: every `rel`-th frame counted back from the last one, frame 0 excluded (routeformer.py:415-419)."""
    return torch.flip(torch.arange(T - 1, 0, -rel).long(), dims=[0])


class StagedBatch(dict):
    """Device-resident batch produced by `Routeformer.stage_batch`: the video tensors hold ONLY the frames the model
    consumes ([B, F, 3, H, W], F = 8 of 40 at the GEM configuration); `video_len[key]` keeps the original length T."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.video_len = {}
        # (event, pinned temporaries) of staging copies that read host memory this object had to pin itself: the raw
        # cudaMemcpy2DAsync of rf_stage_frames_h2d is invisible to PyTorch's caching host allocator, so the temporaries must
        # stay referenced until the copies have run
        self.keepalive = []

    def retire_keepalive(self) -> None:
        for ev, _ in self.keepalive:
            ev.synchronize()
        self.keepalive = []


class Routeformer(nn.Module):
    current_epoch = 0  # LightningModule attribute read by callers / backbones

    def __init__(self, configs: RouteformerConfig, gps_backbone: Optional[Type[nn.Module]] = Informer,
                 video_backbone: Optional[Type[VideoBackboneModule]] = None):
        super().__init__()
        self.configs = configs.copy()
        c = self.configs
        self.with_video = c.with_video if c.with_video is not None else video_backbone is not None
        self.with_scene = c.with_scene
        self.with_gaze = c.with_gaze
        if not self.with_video and self.with_gaze:
            raise ValueError("Current gaze backbone requires a video backbone, but video backbone is not provided.")
        if self.with_video:
            self.video_backbone = video_backbone(configs=c.video_backbone_config)
            E = c.image_embedding_size
            self.frame_encoder = PerceiveEncoder(
                in_channels=self.video_backbone.output_feature_shape[0], out_len=1, out_channels=E, n_heads=c.encoder_heads,
                layers=c.encoder_layers, d_ff=c.encoder_d_ff, dropout=c.feature_dropout)
            self.left_video_embedding = nn.Parameter(torch.randn(1, 1, E))
            self.right_video_embedding = nn.Parameter(torch.randn(1, 1, E))
            self.gaze_video_embedding = nn.Parameter(torch.randn(1, 1, E))
            self.video_output_embedding = nn.Parameter(torch.randn(1, 1, E))
            self.video_encoder = PerceiveEncoder(
                in_channels=E, out_len=c.gps_backbone_config.seq_len, out_channels=c.encoder_hidden_size, n_heads=c.encoder_heads,
                layers=c.encoder_layers, d_ff=c.encoder_d_ff, dropout=c.feature_dropout)
            if self.with_gaze:
                self.gaze_encoder = PerceiveEncoder(
                    in_channels=2, out_len=c.gps_backbone_config.seq_len, out_channels=c.encoder_hidden_size, n_heads=c.encoder_heads,
                    layers=c.encoder_layers, d_ff=c.encoder_d_ff, dropout=c.feature_dropout)
                self.gaze_video_decoder = PerceiveDecoder(
                    query_channels=c.encoder_hidden_size, value_channels=c.encoder_hidden_size, out_channels=c.encoder_hidden_size,
                    out_len=c.gps_backbone_config.seq_len, dropout=c.feature_dropout, d_ff=c.encoder_d_ff,
                    n_heads=c.cross_modal_decoder_heads, layers=c.cross_modal_decoder_layers, mix=False)
        self.gps_backbone = gps_backbone(configs=c.gps_backbone_config)
        self.view_dropout = c.view_dropout
        self.motion_noise = c.motion_noise
        self.gaze_dropout = c.gaze_dropout
        self.feature_dropout = c.feature_dropout
        # test hook: when set to a list, every ProbSparse call appends {"where", "top", "measure"} (the oracle replays them)
        self.record_tops: Optional[list] = None
        # test hook: {module path: [int32 [B,H,u] device tensor per call]} -- ProbSparse selections to use instead of the kernels' own
        self.forced_tops: Optional[dict] = None
        self.last_draw_log: List[tuple] = []
        self._idx_slots = {}
        self._deferred_tables: list = []
        self._pending_plan = None
        # (drop_left, drop_right, drop_gaze) of the most recent draw plan; `forced_pattern` makes the next plans use a given
        # pattern WITHOUT drawing torch.rand (CUDA-graph capture of one graph per drop pattern, parallel.DataParallelTrainer)
        self.last_pattern = (False, False, False)
        self.forced_pattern = None
        # called (no arguments) in the backward pass as soon as the gradient w.r.t. the GPS backbone's input is complete, i.e.
        # when every gradient of the backbone's own parameters has been enqueued (data-parallel trainer: early all-reduce)
        self.backbone_grads_ready_hook = None
        # > 1 only inside `forward_samples`: that many stochastic forwards of one batch run as ONE pass (the index tables of
        # every attention call then hold one table per sample, `idx_group` keeps the samples apart)
        self._samples = 1

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def uses_dropout(self) -> bool:
        """Element-wise dropout anywhere on the path (Perceive feature dropout, GPS-backbone dropout)."""
        gb = float(getattr(self.configs.gps_backbone_config, "dropout", 0.0) or 0.0)
        return self.feature_dropout > 0.0 or gb > 0.0

    # ------------------------------------------------------------------------------------------
    # host staging
    # ------------------------------------------------------------------------------------------
    def stage_batch(self, host_batch: dict, device=None, out: Optional[StagedBatch] = None) -> StagedBatch:
        """Pinned host batch (reference layout: videos [B,T,3,H,W]) -> device batch, moving only the consumed frames.

        A GEM-shaped clip is 41.2 MB of which 8.24 MB (24 frames) are read by the model (routeformer.py:415-421); the
        copies are asynchronous on the current stream."""
        c = self.configs
        device = device or self.device
        reuse = out is not None  # write into the tensors of an existing staged batch (static buffers of a CUDA graph)
        out = out if reuse else StagedBatch()
        out.retire_keepalive()  # temporaries of the previous staging into this object (waits for those copies only)
        temporaries = []
        for key, value in host_batch.items():
            if key.endswith("_video") and self.with_video:
                rel = c.output_fps // (c.gaze_fps if key == "front_video" else c.video_fps)
                T = value.shape[1]
                src = value
                if not (value.is_pinned() and value.is_contiguous()):
                    src = value.contiguous().pin_memory()
                    temporaries.append(src)
                out[key] = ops.stage_frames_h2d(src, frame_indices(T, rel).tolist(), device, out=out[key] if reuse else None)
                out.video_len[key] = T
            elif reuse:
                out[key].copy_(value, non_blocking=True)
            else:
                out[key] = value.to(device, non_blocking=True)
        if temporaries:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            out.keepalive.append((ev, temporaries))
        return out

    def _device_index(self, idx: torch.Tensor, device) -> torch.Tensor:
        """Cached device copy of a small constant index tensor (no H2D copy inside a captured step)."""
        key = ("index", tuple(idx.tolist()), str(device))
        if key not in self._idx_slots:
            self._idx_slots[key] = idx.to(device)
        return self._idx_slots[key]

    @staticmethod
    def _video_len(batch, key) -> int:
        if isinstance(batch, StagedBatch) and key in batch.video_len:
            return batch.video_len[key]
        return batch[key].shape[1]

    # ------------------------------------------------------------------------------------------
    # CPU-RNG plan: every torch.rand / torch.randint of one preprocess+forward, in reference order
    # ------------------------------------------------------------------------------------------
    def _plan_visual(self, batch, training: bool) -> dict:
        c = self.configs
        plan = {"drop_left": False, "drop_right": False, "drop_gaze": False, "entries": {}}
        draws: List[torch.Tensor] = []
        log: List[tuple] = []

        def draw(L_K, L_Q):
            U = sparse_budget(L_K, self.frame_encoder.encoder.attn_layers[0].attention.inner_attention.factor)
            t = torch.randint(L_K, (L_Q, U))
            draws.append(t)
            log.append((L_K, L_Q, U))
            return len(draws) - 1

        n_layers = c.encoder_layers
        S = self.video_backbone.output_feature_shape[1] * self.video_backbone.output_feature_shape[2] + 1
        frame_sets = {}
        n_streams = 1
        if self.with_scene:
            left = batch["left_video"]
            has_right = "right_video" in batch
            drop_left = drop_right = False
            if self.forced_pattern is not None and training:
                drop_left, drop_right = self.forced_pattern[0], self.forced_pattern[1] or not has_right
            elif self.view_dropout > 0.0 and training:  # routeformer.py:405-410
                one = bool(torch.rand(1) < self.view_dropout)
                drop_left = one and bool(torch.rand(1) < 0.5)
                drop_right = (one and not drop_left) or not has_right
            else:
                drop_right = not has_right
            plan["drop_left"], plan["drop_right"] = drop_left and training, drop_right and training
            if not plan["drop_right"]:  # RIGHT is encoded before LEFT (routeformer.py:427-428)
                frame_sets["right"] = [draw(S, S) for _ in range(n_layers)]
            if not plan["drop_left"]:
                frame_sets["left"] = [draw(S, S) for _ in range(n_layers)]
            T_vid = self._video_len(batch, "left_video")
            n_streams += 2
        if self.with_gaze:
            if self.forced_pattern is not None and training:
                plan["drop_gaze"] = bool(self.forced_pattern[2])
            elif self.gaze_dropout > 0.0 and training:  # routeformer.py:300-301
                plan["drop_gaze"] = bool(torch.rand(1) < self.gaze_dropout)
            T_vid = self._video_len(batch, "front_video")
            n_streams += 1
            if not plan["drop_gaze"]:
                frame_sets["front"] = [draw(S, S) for _ in range(n_layers)]
                Lg = c.gps_backbone_config.seq_len
                plan["entries"]["gaze_encoder"] = [draw(Lg, Lg) for _ in range(n_layers)]
                plan["entries"]["gaze_video_decoder"] = [draw(Lg, Lg) for _ in range(c.cross_modal_decoder_layers)]
        Lv = n_streams * T_vid
        plan["entries"]["video_encoder"] = [draw(Lv, Lv) for _ in range(n_layers)]
        plan["frame_sets"], plan["draws"], plan["log"], plan["T_vid"] = frame_sets, draws, log, T_vid
        return plan

    def _plan_backbone(self, T: int, P: int):
        """Informer draws: encoder layers over the distilled lengths, decoder self, decoder cross."""
        gb = self.gps_backbone
        if not isinstance(gb, Informer):
            return [], []
        f = gb.encoder.attn_layers[0].attention.inner_attention.factor
        draws, log = [], []

        def draw(L_K, L_Q):
            U = sparse_budget(L_K, f)
            draws.append(torch.randint(L_K, (L_Q, U)))
            log.append((L_K, L_Q, U))

        L = T
        n = len(gb.encoder.attn_layers)
        for i in range(n):
            draw(L, L)
            if gb.encoder.conv_layers is not None and i < n - 1:
                L = (L + 1) // 2 + 1
        for _ in gb.decoder.layers:
            draw(T + P, T + P)
            draw(L, T + P)
        return draws, log

    def _upload(self, tables: List[torch.Tensor], device, slot: str, refill_only: bool = False) -> List[torch.Tensor]:
        """Index tables -> device views [1, L_Q, U] int32 through a PERSISTENT pinned staging buffer and ONE H2D copy.

        The staging / device buffers are kept per slot (same addresses every step) so that the copy can be captured in a
        CUDA graph; `refill_only` rewrites the pinned buffer on the host without issuing the copy (graph replay does it)."""
        if not tables:
            return []
        n = sum(t.numel() for t in tables)
        key = (slot, n, str(device))
        if key not in self._idx_slots:
            host = torch.empty(n, dtype=torch.int32)
            if device.type == "cuda":
                host = host.pin_memory()
            self._idx_slots[key] = [host, torch.empty(n, dtype=torch.int32, device=device), None]
        host, dev, copied = self._idx_slots[key]
        if refill_only == "defer":  # draw now, write the pinned buffer later (commit_deferred_draws): it may still be in use
            self._deferred_tables.append((host, torch.cat([t.reshape(-1) for t in tables]).to(torch.int32)))
            return []
        capturing = device.type == "cuda" and torch.cuda.is_current_stream_capturing()
        if copied is not None and not capturing:
            # eager path: the previous H2D copy out of this pinned buffer (an earlier forward, the target pass, the previous
            # autoregressive window) may still be queued -- rewriting the buffer now would hand that forward torn index tables.
            # (While a graph is being captured nothing executes: the warm-up passes were synchronised before the capture began.)
            copied.synchronize()
            self._idx_slots[key][2] = None
        off = 0
        for t in tables:
            host[off:off + t.numel()].copy_(t.reshape(-1))
            off += t.numel()
        if refill_only:
            return []
        dev.copy_(host, non_blocking=True)
        if device.type == "cuda" and not capturing:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._idx_slots[key][2] = ev
        out, off = [], 0
        for t in tables:
            out.append(dev[off:off + t.numel()].view(1, *t.shape))
            off += t.numel()
        return out

    def commit_deferred_draws(self) -> bool:
        """Writes the tables drawn by `prepare_draws(refill_only="defer")` into the pinned buffers the captured graph copies from."""
        pending, self._deferred_tables = self._deferred_tables, []
        for host, flat in pending:
            host.copy_(flat)
        return bool(pending)

    def prepare_draws(self, batch, training: bool = None, refill_only=False, backbone: bool = True):
        """Makes every CPU random draw of one (non-autoregressive) forward, in reference order, and stages the index tables.

        Called implicitly by `preprocess_batch` / `_forward`; called explicitly with `refill_only=True` before replaying a
        captured CUDA graph of the step (the graph contains the H2D copy, the host only has to refresh the pinned buffer), or
        with `refill_only="defer"` while the previous replay is still running (the draws are made now, in order, and reach
        the pinned buffer through `commit_deferred_draws` once that replay has consumed it)."""
        if training is None:
            training = self.training
        dev = batch["gps"].device
        plan = {"visual": None, "backbone": None}
        log: List[tuple] = []
        if self.with_video:
            pv = self._plan_visual(batch, training)
            # the training forward and the eval-mode target pass of one step (full_comparison.py:481-482) get separate staging
            # buffers: both uploads of a captured step read theirs at replay time
            pv["tables"] = self._upload(pv["draws"], dev, "visual" if training else "visual_eval", refill_only)
            plan["visual"] = pv
            log += pv["log"]
            self.last_pattern = (bool(pv["drop_left"]), bool(pv["drop_right"]), bool(pv["drop_gaze"]))
        if backbone and isinstance(self.gps_backbone, Informer) and not (not self.training and self.configs.autoregressive):
            draws, blog = self._plan_backbone(batch["gps"].shape[1], self.gps_backbone.pred_len)
            plan["backbone"] = (blog, self._upload(draws, dev, "backbone", refill_only))
            log += blog
        self.last_draw_log = log
        self._pending_plan = None if refill_only else plan
        return plan

    def _source(self, keys_tables, groups=0):
        return PlannedIndexSource([(k, t, groups) for k, t in keys_tables], self.forced_tops)

    def prepare_sample_draws(self, batch, n: int):
        """The CPU draws of `n` consecutive eval forwards of the same batch (full_comparison.py:659-665), in the order those
        forwards would make them (forward 0: visual path then GPS backbone, forward 1: ...), uploaded with one copy.  Returns
        the plan of ONE forward whose every table is stacked [n, L_Q, U]."""
        dev = batch["gps"].device
        visual, backbone, log = [], [], []
        for _ in range(n):
            if self.with_video:
                visual.append(self._plan_visual(batch, False))
                log += visual[-1]["log"]
            if isinstance(self.gps_backbone, Informer):
                backbone.append(self._plan_backbone(batch["gps"].shape[1], self.gps_backbone.pred_len))
                log += backbone[-1][1]

        def stacked(per_sample_draws, slot):
            n_draws = len(per_sample_draws[0])
            views = self._upload([per_sample_draws[s_][i] for i in range(n_draws) for s_ in range(n)], dev, slot)
            # the n tables of draw i are adjacent in the staging buffer: one strided view instead of a concatenation
            return [views[i * n].as_strided((n, *views[i * n].shape[1:]), (views[i * n][0].numel(), views[i * n].shape[2], 1))
                    for i in range(n_draws)]

        plan = {"visual": None, "backbone": None}
        if visual:
            pv = dict(visual[0])
            pv["tables"] = stacked([v["draws"] for v in visual], "visual_samples")
            plan["visual"] = pv
            self.last_pattern = (bool(pv["drop_left"]), bool(pv["drop_right"]), bool(pv["drop_gaze"]))
        if backbone:
            plan["backbone"] = (backbone[0][1], stacked([b[0] for b in backbone], "backbone_samples"))
        self.last_draw_log = log
        return plan

    # ------------------------------------------------------------------------------------------
    # visual streams
    # ------------------------------------------------------------------------------------------
    def _encode_frames(self, batch, plan, dev_tables, training) -> dict:
        """All active camera views through backbone + frame encoder in one batched pass -> {view: [B,F,E]}."""
        c = self.configs
        order = [v for v in ("right", "left", "front") if v in plan["frame_sets"]]
        if not order:
            return {}
        vb = self.video_backbone
        views, n_per_view = [], None
        for name in order:
            if name == "front":
                key, rel = "front_video", c.output_fps // c.gaze_fps
            else:
                key = "left_video" if (name == "left" or "right_video" not in batch) else "right_video"
                rel = c.output_fps // c.video_fps
            assert rel > 0, "Video FPS must be a divisor of the output FPS"
            video = batch[key]
            B = video.shape[0]
            T = self._video_len(batch, key)
            times = frame_indices(T, rel)  # frame times in the original clip
            t_idx = torch.arange(len(times)) if video.shape[1] != T else times  # staged batches are already compact
            centers = None
            if name == "front" and isinstance(vb, PatchEmbedBackbone) and vb.configs.fov == "gaze":
                gaze = batch["gaze"].to(torch.float32)
                g = ops.median_downsample(gaze.contiguous(), T) if gaze.shape[1] > T else gaze
                centers = g[:, self._device_index(times, g.device)].reshape(-1, 2).clamp(0.0, 1.0)
            views.append({"name": name, "video": video, "t_idx": t_idx, "centers": centers, "B": B, "T": T})
            n = B * len(t_idx)
            n_per_view = n if n_per_view is None else n_per_view
            if n != n_per_view:
                raise ValueError("all camera views must provide the same number of encoded frames per clip")
        S = vb.output_feature_shape[1] * vb.output_feature_shape[2] + 1
        if isinstance(vb, PatchEmbedBackbone):
            tokens = vb.encode_views([{**v, "video": v["video"].contiguous()} for v in views])
        else:  # foreign VideoBackboneModule plugin: generic (slower) path of routeformer.py:472-487
            feats = []
            for v in views:
                frames = v["video"][:, self._device_index(v["t_idx"], v["video"].device)].flatten(0, 1)
                f = vb(frames).to(torch.float32)
                f = f.permute(0, 2, 3, 1).reshape(f.shape[0], -1, f.shape[1])
                feats.append(torch.cat([f, -torch.ones_like(f)[:, :1, :]], dim=1))
            tokens = torch.cat(feats, 0).reshape(-1, feats[0].shape[-1]).contiguous()
        n_layers = c.encoder_layers
        keys_tables = []
        ns = self._samples
        for layer in range(n_layers):
            ids = [plan["frame_sets"][name][layer] for name in order]
            if ns == 1:
                table = torch.cat([dev_tables[i] for i in ids], 0).contiguous()
            else:  # sequences are ordered [sample][view][frame]: tables likewise
                table = torch.stack([dev_tables[i] for i in ids], 1).reshape(ns * len(ids), *dev_tables[ids[0]].shape[1:])
            keys_tables.append((plan["log"][ids[0]], table))
        src = self._source(keys_tables, groups=n_per_view if len(order) * ns > 1 else 0)
        n_total = n_per_view * len(order)
        feats = self.frame_encoder.encode(tokens, n_total, S, src, self.record_tops, "frame_encoder", samples=ns)  # [ns*n_total, E]
        E = c.image_embedding_size
        out = {}
        feats = feats.view(ns, len(views), n_per_view, E)
        for i, v in enumerate(views):
            out[v["name"]] = feats[:, i].reshape(ns * v["B"], len(v["t_idx"]), E)
        return out

    overlap_branches = True  # run the gaze encoder concurrently with the frame path (class-level switch, e.g. for debugging)

    def _branch_stream(self, dev) -> torch.cuda.Stream:
        s = getattr(self, "_side_stream", None)
        if s is None or s.device != torch.device(dev):
            s = self._side_stream = torch.cuda.Stream(dev)
        return s

    def _visual_features(self, batch, training: bool, plan, dev_tables) -> torch.Tensor:
        c = self.configs
        E, T = c.image_embedding_size, plan["T_vid"]
        dev = self.device
        rel_v, rel_g = c.output_fps // c.video_fps, c.output_fps // c.gaze_fps
        ns = self._samples
        B = batch["gps"].shape[0] * ns
        groups = batch["gps"].shape[0] if ns > 1 else 0  # sequences per index table (sample-major batch of forward_samples)
        # The gaze encoder (B*40 tokens: ~100 small launches forward, ~200 backward) depends on nothing the frame encoder
        # produces: it is forked onto a side stream BEFORE the frame path is enqueued and joined where the gaze-video decoder
        # needs both, so its launches fill the SMs the big frame-encoder kernels leave idle.  Autograd replays the same stream
        # assignment in the backward pass.
        gq = None
        if self.with_gaze and not plan["drop_gaze"]:
            Lg = c.gps_backbone_config.seq_len
            gaze = batch["gaze"].to(torch.float32).contiguous()
            src = self._source([(plan["log"][i], dev_tables[i]) for i in plan["entries"]["gaze_encoder"]], groups)
            main = torch.cuda.current_stream()
            side = self._branch_stream(dev) if self.overlap_branches else None
            if side is not None:
                side.wait_stream(main)
            with torch.cuda.stream(side if side is not None else main):
                gaze_ds = ops.median_downsample(gaze, Lg)  # utils/filter.py:5-43 (raises if Lg >= samples)
                gaze_in = torch.nn.functional.pad(gaze_ds.view(-1, 2), (0, 2))
                gq = self.gaze_encoder.encode(gaze_in, B // ns, Lg, src, self.record_tops, "gaze_encoder", samples=ns)  # [B*Lg, E]
            if side is not None:
                gaze.record_stream(side)
        feats = self._encode_frames(batch, plan, dev_tables, training)
        streams, srcs, embs = [], [], []

        def frames_stream(name, rel):
            idx = frame_indices(T, rel)
            F_, first, step = len(idx), int(idx[0]), (int(idx[1] - idx[0]) if len(idx) > 1 else 1)
            if name in feats:
                streams.append((True, False, F_, first, step))
                srcs.append(feats[name].contiguous())
            else:  # dropped view: zeros (routeformer.py:464-470)
                streams.append((False, False, F_, first, step))
                srcs.append(None)

        if self.with_scene:
            frames_stream("left", rel_v)
            embs.append(self.left_video_embedding)
            frames_stream("right", rel_v)
            embs.append(self.right_video_embedding)
        if self.with_gaze:
            if plan["drop_gaze"]:
                streams.append((False, True, 0, 0, 1))
                srcs.append(None)
            else:
                idx = frame_indices(T, rel_g)
                meta = dict(B=B, T=T, E=E, streams=[(True, False, len(idx), int(idx[0]), int(idx[1] - idx[0]) if len(idx) > 1 else 1)])
                front_full = Fn.TokenStreams.apply(meta, feats["front"].contiguous(), torch.zeros(E, device=dev))  # [B,T,E]
                ent = plan["entries"]
                if self.overlap_branches:  # join: the decoder reads the gaze-encoder output on the main stream
                    torch.cuda.current_stream().wait_stream(self._branch_stream(dev))
                    gq.record_stream(torch.cuda.current_stream())
                src = self._source([(plan["log"][i], dev_tables[i]) for i in ent["gaze_video_decoder"]], groups)
                g = self.gaze_video_decoder.decode(front_full.view(B * T, E), gq, B, T, Lg, src, self.record_tops, "gaze_video_decoder")
                g = g.view(B, Lg, -1)[:, :T].contiguous()  # routeformer.py:327
                streams.append((True, True, 0, 0, 1))
                srcs.append(g)
            embs.append(self.gaze_video_embedding)
        streams.append((False, True, 0, 0, 1))  # the T "output" tokens carry only their embedding (routeformer.py:340-342)
        srcs.append(None)
        embs.append(self.video_output_embedding)
        meta = dict(B=B, T=T, E=E, streams=streams)
        tokens = Fn.TokenStreams.apply(meta, *srcs, *embs)  # [B, n_streams*T, E]
        n_streams = len(streams)
        src = self._source([(plan["log"][i], dev_tables[i]) for i in plan["entries"]["video_encoder"]], groups)
        vis = self.video_encoder.encode(tokens.view(B * n_streams * T, E), B, n_streams * T, src, self.record_tops, "video_encoder")
        return vis.view(B, -1, c.encoder_hidden_size)

    # ------------------------------------------------------------------------------------------
    # public API
    # ------------------------------------------------------------------------------------------
    def preprocess_batch(self, batch, training: bool = None):
        """-> (motion_dynamics [B,T,2], visual_features [B,T,E] or [])   (routeformer.py:254-348)."""
        if training is None:
            training = self.training
        c = self.configs
        gps = batch["gps"].to(torch.float32)
        if not gps.is_cuda:
            raise RuntimeError("routeformer_b200 runs on CUDA tensors only: move the batch to the GPU (there is no CPU fallback)")
        self.arena = Arena.ensure(self)
        if self.motion_noise > 0.0 and self.training:
            gps = gps + torch.randn_like(gps) * self.motion_noise
        B, T, _ = gps.shape
        x8, _ = Fn.MotionFeatures.apply(gps.contiguous(), None, 0, 8, False, c.normalize_motion, c.motion_mean, c.motion_std, True, False)
        motion = x8[:, :, :2]
        visual: object = []
        # stand-alone calls (e.g. the target pass of full_comparison.py:482) draw only what the reference's preprocess_batch
        # draws; `forward` has already planned the GPS-backbone draws that follow
        plan = self._pending_plan if self._pending_plan is not None else self.prepare_draws(batch, training, backbone=False)
        self._pending_plan = None
        self._backbone_plan = plan["backbone"]
        if self.with_video:
            visual = self._visual_features(batch, training, plan["visual"], plan["visual"]["tables"])
        return motion, visual

    def _forward(self, motion_dynamics, visual_features):
        """-> (raw backbone output [B,P,c_out], origin angles [B])   (routeformer.py:204-252; the rotate-back of :250 is
        folded into postprocess_batch's decode kernel)."""
        c = self.configs
        gb = self.gps_backbone
        B, T, _ = motion_dynamics.shape
        has_vis = self.with_video and not isinstance(visual_features, list)
        E = c.encoder_hidden_size if has_vis else 0
        enc_in = 5 + E
        ld = (enc_in + 3) // 4 * 4
        x, origin = Fn.MotionFeatures.apply(motion_dynamics.contiguous(), visual_features if has_vis else None, E, ld, c.rotate_motion,
                                            False, 0.0, 1.0, bool(c._only_motion) or not has_vis, True)
        if self.backbone_grads_ready_hook is not None and x.requires_grad:
            hook = self.backbone_grads_ready_hook
            x.register_hook(lambda g: (hook(), None)[1])
        if isinstance(gb, Informer):
            if getattr(self, "_backbone_plan", None) is not None:
                log, tables = self._backbone_plan
                self._backbone_plan = None
            else:  # autoregressive windows / direct _forward calls: draw on the spot (same stream order as the reference)
                draws, log = self._plan_backbone(T, gb.pred_len)
                tables = self._upload(draws, x.device, f"backbone_p{gb.pred_len}")
                self.last_draw_log += log
            groups = B // self._samples if self._samples > 1 else 0
            out = gb.run(x, PlannedIndexSource([(k, t, groups) for k, t in zip(log, tables)], self.forced_tops), self.record_tops)
        else:  # foreign GPS backbone plugin (routeformer.py:241)
            out = gb(x[:, :, :enc_in])
        if c.decoder_mode == "recursive":
            if c.dense_prediction and out.shape[-1] != enc_in:
                # routeformer.py:244-245 adds the WHOLE last input row (enc_in columns) to an output of c_out = enc_in - 3 columns:
                # the reference cannot run decoder_mode="recursive" together with dense_prediction, and neither does this
                raise RuntimeError(f"The size of tensor a ({out.shape[-1]}) must match the size of tensor b ({enc_in}) at non-singleton "
                                   "dimension 2 (decoder_mode='recursive' with dense_prediction, routeformer.py:245)")
            out = out + (x[:, -1:, :enc_in] if c.dense_prediction else x[:, -1:, :2])
        return out, origin

    def postprocess_batch(self, last_input_gps, output, origin=None):
        """-> (future_motion_vector, future_gps_positions, future_visual_features)   (routeformer.py:350-395)."""
        c = self.configs
        E = c.image_embedding_size
        rotate = bool(c.rotate_motion) and origin is not None
        wp, motion = Fn.DecodeWaypoints.apply(output.contiguous(), origin, last_input_gps.to(torch.float32).reshape(-1, 2).contiguous(),
                                              rotate, c.normalize_motion, c.motion_mean, c.motion_std)
        wp = wp.to(last_input_gps.dtype)
        rest = output[:, :, 2:]
        dense = None
        if self.with_video and c.dense_prediction:
            assert rest.shape[-1] >= E, f"Output shape for left/right vid. must be at least {E}, but is {rest.shape}."
            dense = rest[:, :, :E]
            rest = rest[:, :, E:]
        assert rest.shape[-1] == 0, f"Output should be empty at this point, but is {rest.shape}."
        return motion, wp, dense

    @torch.no_grad()
    def forward_samples(self, batch, n: int):
        """`n` stochastic eval-mode forwards of one batch as ONE pass -> (waypoints [n,B,P,2], dense [n,B,P,E] or None).

        Equals `[self(batch) for _ in range(n)]` (the loop of full_comparison.py:659-665) draw for draw: the ProbSparse index
        tables are drawn in the order of the sequential forwards and each sample keeps its own tables.  What the samples share
        is computed once (FoV crop, patch embedding, token embedding of the frame encoder, median filters); everything after
        the first sampled attention runs on the n-fold batch.  Autoregressive models keep the sequential loop."""
        c = self.configs
        if self.training:
            raise RuntimeError("forward_samples is an evaluation path: call model.eval() first")
        if c.autoregressive or n == 1:
            outs = [self(batch) for _ in range(n)]
            if c.dense_prediction:
                return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])
            return torch.stack(outs), None
        B = batch["gps"].shape[0]
        self._pending_plan = self.prepare_sample_draws(batch, n)
        self._samples = n
        try:
            motion, visual = self.preprocess_batch(batch)
            motion = motion.repeat(n, 1, 1)
            out, origin = self._forward(motion, visual)
            _, wp, dense = self.postprocess_batch(batch["gps"][:, -1:, :].repeat(n, 1, 1), out, origin)
        finally:
            self._samples = 1
        wp = wp.view(n, B, *wp.shape[1:])
        return wp, (dense.reshape(n, B, *dense.shape[1:]) if dense is not None else None)

    def forward(self, batch, target_batch=None):
        c = self.configs
        if self.training and self.uses_dropout:
            ops.DropoutStream.begin_step(batch["gps"].device)  # fresh Philox sub-streams for this step's dropout sites
        if self._pending_plan is None:
            self.prepare_draws(batch)
        motion, visual = self.preprocess_batch(batch)
        last_gps = batch["gps"][:, -1:, :]
        if not (not self.training and c.autoregressive):
            out, origin = self._forward(motion, visual)
            _, wp, dense = self.postprocess_batch(last_gps, out, origin)
        else:  # autoregressive windows (routeformer.py:164-197)
            if not self.with_video:  # the reference slices its empty visual-feature list at :187
                raise TypeError("list indices must be integers or slices, not tuple")
            outputs, done = [], 0
            step = c.autoregressive_step_size
            pred_len = self.gps_backbone.pred_len
            self.gps_backbone.pred_len = step
            try:
                while done < pred_len:
                    out, origin = self._forward(motion, visual)
                    mv, wp_s, dense_s = self.postprocess_batch(last_gps, out, origin)
                    outputs.append((wp_s, dense_s))
                    motion = torch.cat([motion[:, step:], mv], dim=1)
                    last_gps = wp_s[:, -1:, :]
                    visual = torch.cat([visual[:, step:], dense_s], dim=1)
                    done += step
            finally:
                self.gps_backbone.pred_len = pred_len
            wp = torch.cat([o[0] for o in outputs], dim=1)[:, :pred_len]
            dense = torch.cat([o[1] for o in outputs], dim=1)[:, :pred_len] if self.with_video else None
        if c.dense_prediction:
            return wp, dense
        return wp
