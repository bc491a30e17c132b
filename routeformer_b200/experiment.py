"""The two callers of the hot path in experiments/full_comparison.py, for ONE Routeformer model.

`ParallelTrainerSteps.training_step` = `ParallelTrainer.training_step` (:470-532): forward on the input window, the target pass
`preprocess_batch(target, training=False)` for the dense supervision, the two FutureDiscountedLosses and the dense-loss
re-weighting (`dense_loss_ratio * trajectory_loss / max(dense_loss, 1e-6)`, detached, switched on after epoch 10), plus the
logged ADE / FDE.  `eval_step` = `_eval_step` (:654-679): `torch.manual_seed(12345)`, five stochastic forwards, their mean,
then per-clip loss / ADE / FDE -- here the five forwards run as ONE five-fold batch with per-sample index tables
(`Routeformer.forward_samples`: same CPU draws in the same order; the crop / patch embedding / token embedding they share is
computed once) and the metrics are ONE kernel (`rf_eval_samples`) instead of a Python loop of 3 x B tiny reductions.
The Lightning plumbing around them (logging, PCI buckets, optimiser config) stays with the caller.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import ops
from .metrics import FutureDiscountedLoss, ade, fde


class ParallelTrainerSteps:
    def __init__(self, model, n_eval_samples: int = 5):
        c = model.configs
        self.model = model
        self.n_eval_samples = n_eval_samples
        self.batched_samples = True  # eval_step: the stochastic forwards as one n-fold batch (False: the reference's loop)
        # full_comparison.py:445-454
        self.trajectory_loss = FutureDiscountedLoss(c.discount_factor, c.epsilon, loss_function="smooth_l1")
        self.dense_loss = FutureDiscountedLoss(c.discount_factor, c.visual_epsilon, loss_function="smooth_l1")

    # -- full_comparison.py:470-532 ----------------------------------------------------------------
    def training_step(self, batch: Dict[str, Dict[str, torch.Tensor]], current_epoch: int = 0) -> Tuple[torch.Tensor, dict]:
        model, c = self.model, self.model.configs
        inp, target = batch["train"], batch["target"]
        target_gps = target["gps"].to(torch.float32)
        metrics = {}
        if c.dense_prediction:
            future_gps, future_visual = model(inp)
            with torch.no_grad():
                _, target_visual = model.preprocess_batch(target, training=False)
            target_visual = target_visual[:, : future_visual.shape[1]]
            step = c.autoregressive_step_size
            if c.autoregressive:
                future_gps, target_gps = future_gps[:, :step], target_gps[:, :step]
            trajectory_loss = self.trajectory_loss(future_gps, target_gps)
            if c.autoregressive:
                trajectory_loss = trajectory_loss * (c.gps_backbone_config.pred_len / step)
            target_visual = target_visual.detach()
            if c.autoregressive:
                future_visual, target_visual = future_visual[:, :step], target_visual[:, :step]
            dense_loss = self.dense_loss(future_visual, target_visual)
            weight = (c.dense_loss_ratio * trajectory_loss / torch.clamp(dense_loss, min=1e-6)).detach()
            if current_epoch < 10:  # "Activate dense loss after 10 epochs"
                weight = 0
            metrics["train_dense_loss"] = dense_loss
            loss = trajectory_loss + weight * dense_loss
        else:
            future_gps = model(inp)
            trajectory_loss = self.trajectory_loss(future_gps, target_gps)
            loss = trajectory_loss
        metrics["train_loss"] = trajectory_loss
        metrics["train_ade"] = ade(future_gps, target_gps)
        metrics["train_fde"] = fde(future_gps, target_gps)
        return loss, metrics

    # -- the same step under DataParallelTrainer (CUDA graph, fused AdamW, gradient all-reduce) -------------------
    def graph_hooks(self, current_epoch: int = 0):
        """(step_fn, draw_fn) for `DataParallelTrainer(model, None, step_fn=..., draw_fn=...)`: the whole `training_step`
        (forward, eval-mode target pass, both losses, detached dense re-weighting) becomes ONE captured graph per drop pattern.
        Batches are {"train": {...}, "target": {...}}; set `trainer.graph_tag = current_epoch >= 10` when the epoch crosses the
        point where the dense loss switches on (full_comparison.py:503-506).  The logged metrics of the last step are in
        `self.last_metrics` (static tensors of the captured graph)."""
        model = self.model

        def step_fn(batch, _targets):
            loss, metrics = self.training_step(batch, current_epoch)
            # detached: a metric that kept its autograd graph would keep the previous (warm-up) step's saved tensors alive until
            # the next call replaces it -- i.e. free them in the middle of a graph capture
            self.last_metrics = {k: v.detach() for k, v in metrics.items()}
            return loss

        def draw_fn(batch, refill_only):
            model.prepare_draws(batch["train"], training=True, refill_only=refill_only)
            pattern = model.last_pattern
            if model.configs.dense_prediction and model.with_video:  # the target pass draws only what preprocess_batch draws
                model.prepare_draws(batch["target"], training=False, refill_only=refill_only, backbone=False)
            return pattern

        return step_fn, draw_fn

    # -- full_comparison.py:654-679 ----------------------------------------------------------------
    @torch.no_grad()
    def eval_step(self, batch: Dict[str, Dict[str, torch.Tensor]]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Returns per-clip (losses, ades, fdes), each [B].  The model must be in eval mode, as under Lightning's validation loop."""
        model = self.model
        torch.manual_seed(12345)
        inp, target_gps = batch["train"], batch["target"]["gps"]
        if self.batched_samples and hasattr(model, "forward_samples"):
            preds = model.forward_samples(inp, self.n_eval_samples)[0]  # [n, B, P, 2]: the n forwards as one pass, same draws
        else:
            preds = []
            for _ in range(self.n_eval_samples):
                out = model(inp)
                preds.append(out[0] if model.configs.dense_prediction else out)
            preds = torch.stack(preds)
        loss = self.trajectory_loss
        if loss.current_epoch in loss.discount_factor_dict:
            loss.current_discount_factor = loss.discount_factor_dict[loss.current_epoch]
        eps = 0.0 if loss.epsilon is None else float(loss.epsilon)
        self.last_mean_prediction, per_clip = ops.eval_samples(preds.contiguous(), target_gps, float(loss.current_discount_factor), eps,
                                                               loss.loss_function)
        torch.seed()
        return per_clip[:, 0], per_clip[:, 1], per_clip[:, 2]
