"""Shared helpers for the parity tests (test infrastructure)."""
import os

import torch

from oracle import routeformer_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def case_from_golden(gold):
    """(cfg, spec, state_dict, batch) regenerated from the seeds stored in a golden file."""
    cfg = O.OracleConfig(**gold["cfg"])
    spec = O.BackboneSpec(**gold["spec"]) if gold["spec"] else None
    sd = O.fill_state_dict(O.state_dict_template(cfg, spec), gold["wseed"])
    batch = O.synthetic_batch(gold["B"], cfg, gold["shapes"], seed=gold["dseed"])
    return cfg, spec, sd, batch


def targets_for(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, cfg.pred_len, 2, generator=g) * 5.0 + 80.0,
            torch.randn(B, cfg.pred_len, cfg.image_embedding_size, generator=g))


def rel_err(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def build_product(cfg, spec, fov="frame", informer_dropout: float = 0.0, **overrides):
    """routeformer_b200.Routeformer for an OracleConfig / BackboneSpec pair."""
    import routeformer_b200 as R

    g = R.GPSBackboneConfig(seq_len=cfg.seq_len, label_len=cfg.seq_len, pred_len=cfg.pred_len, factor=cfg.factor, distil=cfg.distil,
                            dropout=informer_dropout, activation=cfg.activation, d_model=cfg.d_model, n_heads=cfg.n_heads, e_layers=cfg.e_layers,
                            d_layers=cfg.d_layers, d_ff=cfg.d_ff)
    vb = None
    if spec is not None:
        vb = R.PatchBackboneConfig(image_size=spec.image_size, patch=spec.patch, channels=spec.channels, fov=fov, window=spec.window)
    kw = dict(gps_backbone_config=g, video_backbone_config=vb, decoder_mode=cfg.decoder_mode, with_video=cfg.with_video,
              with_scene=cfg.with_scene, with_gaze=cfg.with_gaze, dense_prediction=cfg.dense_prediction,
              image_embedding_size=cfg.image_embedding_size, encoder_hidden_size=cfg.encoder_hidden_size,
              encoder_heads=cfg.encoder_heads, encoder_layers=cfg.encoder_layers, encoder_d_ff=cfg.encoder_d_ff,
              cross_modal_decoder_heads=cfg.cross_modal_decoder_heads, cross_modal_decoder_layers=cfg.cross_modal_decoder_layers,
              rotate_motion=cfg.rotate_motion, normalize_motion=cfg.normalize_motion, motion_mean=cfg.motion_mean,
              motion_std=cfg.motion_std, video_fps=cfg.video_fps, gaze_fps=cfg.gaze_fps, output_fps=cfg.output_fps,
              view_dropout=cfg.view_dropout, gaze_dropout=cfg.gaze_dropout, autoregressive=cfg.autoregressive,
              autoregressive_step_size=cfg.autoregressive_step_size)
    kw.update(overrides)
    rc = R.RouteformerConfig(**kw)
    return R.Routeformer(rc, gps_backbone=R.Informer, video_backbone=R.PatchEmbedBackbone if spec is not None else None)


def to_device(batch, device):
    return {k: v.to(device) for k, v in batch.items()}


def tops_for_oracle(recorded, view_order):
    """Product-recorded ProbSparse selections -> {module path: [top, ...]} in the ORACLE's call order.

    The product encodes the camera views in one batched frame-encoder pass (`view_order`, e.g. [right, left, front]);
    the oracle (like the reference) calls the frame encoder once per view in that same order."""
    out = {}
    for rec in recorded:
        top = rec["top"].cpu().long()
        if rec["where"].startswith("frame_encoder") and len(view_order) > 1:
            out.setdefault(rec["where"], []).extend(list(top.chunk(len(view_order), dim=0)))
        else:
            out.setdefault(rec["where"], []).append(top)
    return out


# TF32-derived near-tie bound for the ProbSparse sparsity measure M = max_j s - sum_j s / L_K, s = q.k:
# q and k leave a tcgen05 kind::tf32 GEMM whose operands carry a unit round-off of 2^-11 each, i.e. ~2^-10 relative on every
# product, and the activations feeding it have crossed several such layers.  A pick is "explained" when it lies within
# rel_tol of the u-th largest measure, relative to the largest |M| of that (b, h) problem:
#   Perceive* modules (d_model 128, LayerNorm-ed activations, K = 128 reductions):            2^-7  = 8 x 2^-10
#   Informer (d_model 832: K = 832..3328 reductions, dh = 104 score sums, un-normalised
#             68-channel input, autoregressive windows feed predictions back in):            2^-6  = 16 x 2^-10
# Round 1 used 2e-2 everywhere; the worst gap actually observed per module family is logged by the tests (measured in round
# 2 over 7.4 M picks: Perceive* 5.9e-3, Informer 1.07e-2).
SELECTION_REL_TOL = {"default": 2.0 ** -7, "gps_backbone": 2.0 ** -6}


def _sel_tol(where: str, rel_tol):
    if not isinstance(rel_tol, dict):
        return rel_tol
    for prefix, tol in rel_tol.items():
        if prefix != "default" and where.startswith(prefix):
            return tol
    return rel_tol["default"]


def selection_violations(recorded, oracle_tops, view_order, rel_tol=SELECTION_REL_TOL, stats=None):
    """Counts product selections that are NOT explained by a near-tie in the oracle's sparsity measure.
    `stats` (dict, optional) receives "worst_gap": the largest distance of a mismatching pick from the u-th measure, in
    units of max|M| (also per module family under "worst_by_module"), and "mismatches": picks outside the oracle's own set."""
    import torch

    bad = total = mism = 0
    worst = 0.0
    by_module = {}
    queues = {}
    for t in oracle_tops:
        queues.setdefault(t["where"], []).append(t["measure"])
    for rec in recorded:
        ms = queues[rec["where"]]
        if rec["where"].startswith("frame_encoder") and len(view_order) > 1:
            m = torch.cat([ms.pop(0) for _ in view_order], 0)
        else:
            m = ms.pop(0)
        top = rec["top"].cpu().long()
        u = top.shape[-1]
        kth = m.topk(u, dim=-1).values[..., -1:]                    # u-th largest measure per (b,h)
        scale = m.abs().amax(dim=-1, keepdim=True).clamp_min(1.0)
        tol = _sel_tol(rec["where"], rel_tol) * scale
        sel = torch.zeros_like(m, dtype=torch.bool).scatter(-1, top, True)
        bad += int(((m < kth - tol) & sel).sum() + ((m > kth + tol) & ~sel).sum())
        wrong = ((m < kth) & sel) | ((m > kth) & ~sel)
        mism += int(wrong.sum())
        if wrong.any():
            w = float((((m - kth).abs() / scale)[wrong]).max())
            worst = max(worst, w)
            fam = rec["where"].split(".")[0]
            by_module[fam] = max(by_module.get(fam, 0.0), w)
        total += sel.numel()
    if stats is not None:
        stats["worst_gap"] = max(stats.get("worst_gap", 0.0), worst)
        stats["mismatches"] = stats.get("mismatches", 0) + mism
        for fam, w in by_module.items():
            stats.setdefault("worst_by_module", {})[fam] = max(stats.get("worst_by_module", {}).get(fam, 0.0), w)
    return bad, total


def flips_vs_oracle(recorded, oracle_tops, view_order) -> int:
    """Number of (b, h) ProbSparse problems in which the product selected a different top-u SET than an oracle run on the same
    CPU draws (oracle_tops = the `.tops` of an un-replayed oracle forward, i.e. the reference's own selections)."""
    mine = tops_for_oracle(recorded, view_order)
    n = 0
    for t in oracle_tops:
        got = mine[t["where"]].pop(0)
        n += int((got.sort(-1).values != t["top"].long().sort(-1).values).any(-1).sum())
    return n


def tops_for_product(oracle_tops, view_order, device="cuda"):
    """Oracle selections (`.tops` of an un-replayed oracle forward = the reference's own) -> {module path: [int32 [B,H,u], ...]}
    in the PRODUCT's call order (Routeformer.forced_tops): the per-view frame-encoder calls of the oracle are concatenated along
    the batch axis in view order, because the product encodes all views in one batched pass."""
    grouped = {}
    for t in oracle_tops:
        grouped.setdefault(t["where"], []).append(t["top"])
    out = {}
    for where, tops in grouped.items():
        if where.startswith("frame_encoder") and len(view_order) > 1:
            k = len(view_order)
            assert len(tops) % k == 0
            tops = [torch.cat(tops[i:i + k], 0) for i in range(0, len(tops), k)]
        out[where] = [t.to(torch.int32).contiguous().to(device) for t in tops]
    return out


def first_flip_gap(recorded, oracle_tops, view_order) -> float:
    """Distance (in units of max|M|) between the first differing pick of a product run and the oracle's u-th measure, taken at
    the FIRST ProbSparse call whose selection differs: up to that call both runs saw the same inputs, so the oracle's measure is
    directly comparable.  A value at fp32 rounding level means the flip is a genuine tie no implementation can be expected to
    break the same way; 0.0 if nothing flipped."""
    mine = tops_for_oracle(recorded, view_order)
    for t in oracle_tops:
        got = mine[t["where"]].pop(0)
        m = t["measure"]
        differs = (got.sort(-1).values != t["top"].long().sort(-1).values).any(-1)
        if differs.any():
            u = got.shape[-1]
            kth = m.topk(u, dim=-1).values[..., -1:]
            scale = m.abs().amax(dim=-1, keepdim=True).clamp_min(1.0)
            sel = torch.zeros_like(m, dtype=torch.bool).scatter(-1, got, True)
            wrong = ((m < kth) & sel) | ((m > kth) & ~sel)
            return float((((m - kth).abs() / scale)[wrong]).max()) if wrong.any() else 0.0
    return 0.0


def same_selections(rec_a, rec_b) -> bool:
    """True if two recorded runs picked the same top-u SETS in every ProbSparse call."""
    if len(rec_a) != len(rec_b):
        return False
    for a, b in zip(rec_a, rec_b):
        if a["where"] != b["where"] or not torch.equal(a["top"].sort(-1).values, b["top"].sort(-1).values):
            return False
    return True


def log_parity(line: str, path: str = "gpurun_out/parity_raw_errors.txt") -> None:
    """Raw (no-replay) errors against the reference goldens: printed (pytest -s / -rP) and appended to a file that travels back."""
    print(line)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


class ReplayDraw:
    """Oracle index source: the reference's CPU randint stream + the product's recorded top-u selections."""

    def __init__(self, tops):
        self.draws = O.CpuRandint()
        self.tops = tops

    def __call__(self, L_K, L_Q, U):
        return self.draws(L_K, L_Q, U)

    def next_top(self, where):
        return self.tops[where].pop(0).long().cpu()


def condition_weights(sd):
    """Same seed-determined weights, rescaled so that attention soft-maxes are not saturated.  With the raw N(0,1/fan_in)
    weights the reference algorithm's own gradients move by ~11% (median) under TF32 operand rounding and by ~95% in the
    first Informer attention layer (tools/tf32_sensitivity.py); on this variant TF32 moves them by <= ~3%."""
    out = {}
    for k, v in sd.items():
        if k.endswith("query_projection.weight") or k.endswith("key_projection.weight"):
            v = v * 0.3
        elif "temporal_embedding" in k:
            v = v * 0.02
        elif k.endswith("tokenConv.weight") and k.startswith("gps_backbone"):
            v = v * 0.3
        out[k] = v
    return out


class ReplayDropout:
    """Oracle dropout hook that replays the CUDA path's Philox masks (oracle.dropout_hook).

    `log` = routeformer_b200.ops.DropoutStream.log of the product run: (where, seed, offset, rows, cols) per call site, in call
    order.  The product encodes all camera views in one batched frame-encoder pass, the oracle one view at a time: a product
    record with k times the oracle's rows is consumed by k consecutive oracle calls of that site."""

    def __init__(self, log, p: float, device="cuda"):
        self.p, self.device = p, device
        self.sites = {}
        for where, seed, offset, rows, cols in log:
            self.sites.setdefault(where, []).append([seed, offset, rows, cols, None, 0])
        self.used = 0

    def __call__(self, x, where):
        import torch

        from routeformer_b200 import ops

        rec = self.sites[where][0]
        seed, offset, rows, cols, mask, taken = rec
        if mask is None:
            ones = torch.ones(rows, cols, device=self.device)
            mask = rec[4] = ops.dropout(ones, torch.empty_like(ones), self.p, seed, offset).cpu()
        assert x.shape[-1] == cols, (where, tuple(x.shape), rows, cols)
        n = x.numel() // cols
        out = x * mask[taken:taken + n].reshape(x.shape)
        rec[5] = taken + n
        assert rec[5] <= rows, where
        if rec[5] == rows:
            self.sites[where].pop(0)
        self.used += 1
        return out

    def exhausted(self) -> bool:
        return all(len(v) == 0 for v in self.sites.values())
