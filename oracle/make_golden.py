"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Each golden holds only small tensors: the reference's outputs (and, for the training case,
loss / gradients / BatchNorm running-stat updates) for seed-determined weights
(`fill_state_dict`) and seed-determined synthetic batches (`synthetic_batch`), which the tests
regenerate from the seeds.  The reference's `state_dict` layout (key, shape, dtype) and the
ordered list of its CPU `torch.randint` draws are stored as additional pins.
"""
from __future__ import annotations

import dataclasses
import os
import sys

import torch

from oracle import reference_shim as R
from oracle import routeformer_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL = dict(d_model=64, n_heads=4, e_layers=3, d_ff=128, with_video=True, with_gaze=True, dense_prediction=True,
             encoder_layers=2, encoder_d_ff=64, image_embedding_size=32, encoder_hidden_size=32)
SMALL_SPEC = dict(image_size=32, patch=8, channels=48)
PAPER_FULL = dict(with_video=True, with_gaze=True, dense_prediction=True)
PAPER_SPEC = dict(image_size=256, patch=32, channels=1024)

CASES = {
    # name: (cfg kwargs, spec kwargs or None, shapes, B, weight seed, data seed, mode)
    "gps_only_paper": (dict(), None, "gem", 8, 0, 1, "eval"),  # BASELINE.json configs[0]
    "full_small_eval": (SMALL, SMALL_SPEC, "tiny", 3, 7, 1, "eval"),
    "full_small_train": (SMALL, SMALL_SPEC, "tiny", 3, 7, 2, "train"),
    "full_paper_eval": (PAPER_FULL, PAPER_SPEC, "gem", 2, 3, 4, "eval"),  # configs[1] at B=2
    "dreyeve_small": ({**SMALL, "rotate_motion": True}, SMALL_SPEC, "tiny", 2, 5, 6, "eval"),
    "normalized_small": ({**SMALL, "normalize_motion": True, "motion_mean": 1.83, "motion_std": 0.91},
                         SMALL_SPEC, "tiny", 2, 5, 7, "eval"),
    "no_gaze_small": ({**SMALL, "with_gaze": False}, SMALL_SPEC, "tiny", 2, 8, 9, "eval"),
    "no_scene_small": ({**SMALL, "with_scene": False}, SMALL_SPEC, "tiny", 2, 8, 10, "eval"),
    # routeformer.py:164-197: 30 predictions in windows of 8 (4 windows, the last one truncated) and of 15 with rotation.
    # GPS-only autoregression is not a case: the reference slices its empty visual-feature LIST there and raises TypeError (:187)
    "autoregressive_small": ({**SMALL, "autoregressive": True, "autoregressive_step_size": 8}, SMALL_SPEC, "tiny", 2, 13, 14, "eval"),
    "autoregressive_dreyeve_small": ({**SMALL, "rotate_motion": True, "autoregressive": True, "autoregressive_step_size": 15},
                                     SMALL_SPEC, "tiny", 3, 15, 16, "eval"),
    "sparse_small": ({**SMALL, "dense_prediction": False, "decoder_mode": "vanilla"}, SMALL_SPEC, "tiny", 2, 11, 12, "eval"),
    # round 2: the configurations BASELINE.json benches, at their own sizes
    "full_paper_train": (PAPER_FULL, PAPER_SPEC, "gem", 8, 3, 5, "train"),        # configs[2] per-GPU shard shape, B = 8
    "full_paper_b64_eval": (PAPER_FULL, PAPER_SPEC, "gem", 64, 3, 6, "eval"),     # configs[1]: batch 64 forward
    "dreyeve_paper_eval": ({**PAPER_FULL, "rotate_motion": True}, PAPER_SPEC, "dreyeve", 2, 3, 7, "eval"),  # configs[3] shapes
}
# gradients of tensors up to this many elements are stored in full (the paper-size model keeps the fixture small)
GRAD_SMALL_MAX = {"full_paper_train": 1024}


class DrawLog:
    """Records the argument tuples of every CPU torch.randint issued by the reference."""

    def __enter__(self):
        self.log = []
        self._orig = torch.randint

        def wrapped(*a, **k):
            self.log.append((int(a[0]), tuple(int(x) for x in a[1])))
            return self._orig(*a, **k)

        torch.randint = wrapped
        return self

    def __exit__(self, *exc):
        torch.randint = self._orig


def targets_for(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, cfg.pred_len, 2, generator=g) * 5.0 + 80.0,
            torch.randn(B, cfg.pred_len, cfg.image_embedding_size, generator=g))


def generate(name):
    ckw, skw, shapes, B, wseed, dseed, mode = CASES[name]
    ref = R.load()
    cfg = O.OracleConfig(**ckw)
    spec = O.BackboneSpec(**skw) if skw else None
    model = R.build_reference_model(cfg, spec)
    layout = [(k, tuple(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]
    sd = O.fill_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd)
    batch = O.synthetic_batch(B, cfg, shapes, seed=dseed)
    gold = {"cfg": ckw, "spec": skw, "shapes": shapes, "B": B, "wseed": wseed, "dseed": dseed, "mode": mode,
            "layout": layout}
    model.train(mode == "train")
    torch.manual_seed(12345)  # mirrors _eval_step (experiments/full_comparison.py:655)
    with DrawLog() as dl:
        if mode == "eval":
            with torch.no_grad():
                out = model(batch)
        else:
            out = model(batch)
    gold["draws"] = dl.log
    wp, dense = out if isinstance(out, tuple) else (out, None)
    gold["waypoints"] = wp.detach().clone()
    if dense is not None:
        gold["dense"] = dense.detach().clone()
    t_wp, t_dense = targets_for(cfg, B, dseed + 1000)
    gold["ade"] = ref.ade(wp.detach(), t_wp).item()
    gold["fde"] = ref.fde(wp.detach()[-1:], t_wp[-1:]).item()  # per-sample call, full_comparison.py:667-674
    gold["fde_batch"] = ref.fde(wp.detach(), t_wp).item()
    if mode == "train":
        lossf = ref.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
        loss = lossf(wp, t_wp) + 0.5 * lossf(dense, t_dense)
        loss.backward()
        gold["loss"] = loss.item()
        gold["grad_norm"] = {k: p.grad.norm().item() for k, p in model.named_parameters() if p.grad is not None}
        gold["grad_none"] = [k for k, p in model.named_parameters() if p.grad is None]
        gold["grad_small"] = {k: p.grad.clone() for k, p in model.named_parameters()
                              if p.grad is not None and p.numel() <= GRAD_SMALL_MAX.get(name, 8192)}
        new_sd = model.state_dict()
        gold["bn"] = {k: new_sd[k].clone() for k in new_sd if "running_" in k or "num_batches" in k}
    os.makedirs(OUT, exist_ok=True)
    torch.save(gold, os.path.join(OUT, name + ".pt"))
    print(name, "waypoints", tuple(wp.shape), "draws", len(dl.log))


def generate_submodules():
    """Golden vectors for the blocks below the Routeformer glue + metrics + median filter."""
    ref = R.load()
    gold = {}
    g = torch.Generator().manual_seed(99)
    cfg = O.OracleConfig(encoder_layers=2, encoder_d_ff=64, cross_modal_decoder_layers=2, cross_modal_decoder_heads=4)
    # PerceiveEncoder, frame-like: 9 tokens of 40 channels -> last token
    enc = ref.PerceiveEncoder(in_channels=40, out_len=1, out_channels=24, n_heads=8, layers=2, d_ff=64, dropout=0.0).eval()
    enc.load_state_dict(O.fill_state_dict(enc.state_dict(), 21))
    x = torch.randn(5, 9, 40, generator=g)
    torch.manual_seed(3)
    with torch.no_grad():
        gold["perceive_encoder"] = {"x": x, "y": enc(x)}
    # PerceiveDecoder (mix=False as in routeformer.py:114)
    dec = ref.PerceiveDecoder(query_channels=16, value_channels=16, out_channels=16, out_len=12, dropout=0.0, d_ff=32,
                              n_heads=4, layers=2, mix=False).eval()
    dec.load_state_dict(O.fill_state_dict(dec.state_dict(), 22))
    xe, xd = torch.randn(3, 10, 16, generator=g), torch.randn(3, 12, 16, generator=g)
    torch.manual_seed(4)
    with torch.no_grad():
        gold["perceive_decoder"] = {"x_enc": xe, "x_dec": xd, "y": dec(xe, xd)}
    # median filter
    z = torch.randn(4, 83, 2, generator=g)
    gold["median"] = {"x": z, "y": ref.median_downsampler(z, 20), "y2": ref.median_downsampler(z[:, :80], 40)}
    # metrics / loss
    p, t = torch.randn(6, 30, 2, generator=g), torch.randn(6, 30, 2, generator=g)
    lossf = ref.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
    gold["metrics"] = {"p": p, "t": t, "ade": ref.ade(p, t), "fde": ref.fde(p, t), "fde_1": ref.fde(p[2:3], t[2:3]),
                       "loss": lossf(p * 3, t), "const_offset_fde": ref.fde(torch.zeros(1, 30, 2) + 0.5, torch.zeros(1, 30, 2))}
    for kind in ("mse", "mae"):
        gold["metrics"]["loss_" + kind] = ref.FutureDiscountedLoss(0.9, epsilon=0.3, loss_function=kind)(p * 3, t)
    torch.save(gold, os.path.join(OUT, "submodules.pt"))
    print("submodules ok")


def generate_steps():
    """The two callers of the path, experiments/full_comparison.py:470-532 (training_step) and :654-679 (_eval_step), evaluated
    with the reference's own model / loss / metric objects.  The LightningModule that hosts them needs lightning, wandb and
    the datasets, so the few lines of step logic are replayed here verbatim around the reference components."""
    ref = R.load()
    cfg, spec = O.OracleConfig(**SMALL), O.BackboneSpec(**SMALL_SPEC)
    B, wseed, dseed = 3, 17, 18
    model = R.build_reference_model(cfg, spec)
    sd = O.fill_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd)
    batch = {"train": O.synthetic_batch(B, cfg, "tiny", seed=dseed), "target": O.synthetic_batch(B, cfg, "tiny", seed=dseed + 1, T=cfg.pred_len)}
    # continue the target track from the end of the input track, as a real clip does
    batch["target"]["gps"] = batch["target"]["gps"] + batch["train"]["gps"][:, -1:]
    discount, eps, veps, ratio = {0: 0.97}, 1.0, 0.3, 0.5
    traj_loss = ref.FutureDiscountedLoss(discount, eps, loss_function="smooth_l1")
    dense_lossf = ref.FutureDiscountedLoss(discount, veps, loss_function="smooth_l1")
    gold = {"cfg": SMALL, "spec": SMALL_SPEC, "B": B, "wseed": wseed, "dseed": dseed, "discount": discount, "epsilon": eps,
            "visual_epsilon": veps, "dense_loss_ratio": ratio}
    # ---- training_step (:470-532), epochs 0 (dense weight 0) and 10 (weight active)
    model.train()
    for epoch in (0, 10):
        model.zero_grad()
        torch.manual_seed(12345)
        with DrawLog() as dl:
            inp, target = batch["train"], batch["target"]
            target_gps = target["gps"].to(torch.float32)
            future_gps, future_visual_features = model(inp)
            _, target_visual_features = model.preprocess_batch(target, training=False)
            target_visual_features = target_visual_features[:, : future_visual_features.shape[1]]
            trajectory_loss = traj_loss(future_gps, target_gps)
            target_visual_features = target_visual_features.detach()
            dense_loss = dense_lossf(future_visual_features, target_visual_features)
            dense_loss_weight = (ratio * trajectory_loss / max(dense_loss, 1e-6)).detach()
            if epoch < 10:
                dense_loss_weight = 0
            loss = trajectory_loss + dense_loss_weight * dense_loss
        loss.backward()
        gold[f"train_epoch{epoch}"] = {
            "loss": loss.item(), "trajectory_loss": trajectory_loss.item(), "dense_loss": dense_loss.item(),
            "ade": ref.ade(future_gps, target_gps).item(), "fde": ref.fde(future_gps, target_gps).item(),
            "target_visual": target_visual_features.clone(), "draws": dl.log,
            "grad_norm": {k: p.grad.norm().item() for k, p in model.named_parameters() if p.grad is not None},
        }
    # ---- _eval_step (:654-679), from the same initial state (the training passes above moved the BatchNorm running statistics)
    model.load_state_dict(sd)
    model.eval()
    torch.manual_seed(12345)
    with torch.no_grad(), DrawLog() as dl:
        inter = []
        for _ in range(5):
            future_gps, _ = model(batch["train"])
            inter.append(future_gps)
        future_gps = torch.stack(inter).mean(dim=0)
        losses, ades, fdes = [], [], []
        target_gps = batch["target"]["gps"]
        for index in range(future_gps.shape[0]):
            fgps, tgps = future_gps[index:index + 1], target_gps[index:index + 1]
            losses.append(traj_loss(fgps, tgps))
            ades.append(ref.ade(fgps, tgps))
            fdes.append(ref.fde(fgps, tgps))
    gold["eval"] = {"mean_prediction": future_gps.clone(), "samples": torch.stack(inter), "losses": torch.stack(losses),
                    "ades": torch.stack(ades), "fdes": torch.stack(fdes), "draws": dl.log}
    torch.save(gold, os.path.join(OUT, "steps_small.pt"))
    print("steps ok: train loss", gold["train_epoch0"]["loss"], gold["train_epoch10"]["loss"], "eval ade", gold["eval"]["ades"].tolist())


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    names = sys.argv[1:] or list(CASES)
    if names == ["steps"]:
        generate_steps()
        names = []
    for n in names:
        generate(n)
    if len(sys.argv) == 1:
        generate_submodules()
        generate_steps()
