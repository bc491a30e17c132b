"""GPU end-to-end parity: routeformer_b200.Routeformer (CUDA, through the C ABI) vs the CPU oracle and the golden
vectors generated from the unmodified reference.

ProbSparse top-u selection is a discontinuous function of the scores: at a near-tie the TF32 GEMMs upstream can
legitimately pick a different query than fp32 CPU arithmetic (SURVEY 7 hard part 2).  So every case is checked twice:
  (a) TIGHT: the oracle replays the product's selections; outputs must then agree to TF32 accuracy, and every product
      selection must be explained by the oracle's own sparsity measure up to a small tolerance (no unexplained picks);
  (b) RAW: against the golden vector of the reference with no replay, at the looser "one marginal query flipped" level.
Tolerances (fp32 accumulate, TF32 operands): waypoints 1e-3 relative (the north-star contract); dense features 1e-2.
"""
import pytest
import torch

from oracle import routeformer_oracle as O
from tests.helpers import (ReplayDraw, ReplayDropout, build_product, case_from_golden, first_flip_gap, flips_vs_oracle, load_golden, log_parity,
                           rel_err, same_selections, selection_violations, targets_for, to_device, tops_for_oracle, tops_for_product)

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = ["gps_only_paper", "full_small_eval", "dreyeve_small", "normalized_small", "no_gaze_small", "no_scene_small", "sparse_small",
         "full_paper_eval", "autoregressive_small", "autoregressive_dreyeve_small",
         "dreyeve_paper_eval",     # BASELINE configs[3]: real DR(eye)VE frame shapes 216x768 / 216x384 / 240x320, rotate_motion
         "full_paper_b64_eval"]    # BASELINE configs[1]: paper configuration, batch 64


def view_order(cfg):
    return (["right", "left"] if cfg.with_scene and cfg.with_video else []) + (["front"] if cfg.with_gaze else [])


def _forward(model, batch, precise: bool, forced=None):
    from routeformer_b200 import ops

    model.record_tops = []
    model.forced_tops = forced
    torch.manual_seed(12345)
    try:
        with torch.no_grad(), ops.precise(precise):
            out = model(batch)
        torch.cuda.synchronize()
    finally:
        model.forced_tops = None
    wp, dense = out if isinstance(out, tuple) else (out, None)
    return wp, dense, model.record_tops


def metric_tol(v):
    """"ADE/FDE identical to 4 decimal places": 5e-5 absolute, plus what fp32 positions allow -- the waypoints are fp32 numbers
    of magnitude 80..1000 m (ulp 8e-6..6e-5 m) accumulated over a 30-step cumsum, so a metric of magnitude v reproduces to
    ~1e-5 relative at best (the CPU oracle itself agrees with the reference to 2e-6 relative on the waypoints)."""
    return 5e-5 + 1e-5 * abs(v)


@pytest.mark.parametrize("name", CASES)
def test_eval_forward(name):
    import routeformer_b200 as R

    gold = load_golden(name)
    cfg, spec, sd, batch = case_from_golden(gold)
    model = build_product(cfg, spec).to(DEV).eval()
    model.load_state_dict(sd)
    dev_batch = to_device(batch, DEV)
    wp, dense, tops = _forward(model, dev_batch, precise=False)
    # the CPU RNG stream is consumed exactly like the reference does
    assert [(lk, (lq, u)) for lk, lq, u in model.last_draw_log] == [tuple(d) for d in gold["draws"]]
    # (a) oracle with the product's selections replayed
    torch.manual_seed(12345)
    orc = O.Routeformer(sd, cfg, spec)
    Draw = ReplayDraw(tops_for_oracle(tops, view_order(cfg)))

    with torch.no_grad():
        ref = orc.forward(batch, training=False, draw=Draw)
    ref_wp, ref_dense = ref if isinstance(ref, tuple) else (ref, None)
    assert rel_err(wp.cpu(), ref_wp) < 1e-3
    disp, ref_disp = wp.cpu() - batch["gps"][:, -1:], ref_wp - batch["gps"][:, -1:]
    assert rel_err(disp, ref_disp) < 1e-2
    if dense is not None:
        assert rel_err(dense.cpu(), ref_dense) < 1e-2
    stats = {}
    bad, total = selection_violations(tops, orc.tops, view_order(cfg), stats=stats)
    assert bad == 0, f"{bad} of {total} top-u selections not explained by a TF32 near-tie (worst gaps {stats.get('worst_by_module')})"
    # (b) against the reference's golden output.  North-star bound: 1e-3 relative on the waypoints, ADE / FDE to 4 decimals.
    # The oracle run WITHOUT replay reproduces the golden (asserted in the CPU suite), so its selections are the reference's own.
    #   (i)   precise mode (3xTF32 GEMMs, fp32-level) with the reference's selections FORCED into the kernels: pure arithmetic
    #         parity -- waypoints <= 2e-5 relative and ADE / FDE of the GPU prediction equal to the reference's values to 4
    #         decimal places (fp32-limited, see metric_tol), asserted for every case;
    #   (ii)  precise mode, own selections (RAW, no replay): <= 1e-3; if a query flipped, the first flip must be a tie at fp32
    #         rounding level (gap <= 1e-4 of max|M|: summation order decides it, no implementation can reproduce it), and the
    #         bound is the "one marginal query flipped" level 5e-3;
    #   (iii) default TF32 mode, own selections (RAW): <= 1e-3 when no query flipped, else <= 5e-3 -- a flipped marginal query
    #         changes which rows get real attention, an O(1) local effect the reference itself shows under TF32 operand rounding.
    # `flips` counts the (b, h) problems in which the GPU picked a different top-u set than the reference.
    torch.manual_seed(12345)
    orc_raw = O.Routeformer(sd, cfg, spec)
    with torch.no_grad():
        orc_raw.forward(batch, training=False)
    vo = view_order(cfg)
    wp_f, dense_f, _ = _forward(model, dev_batch, precise=True, forced=tops_for_product(orc_raw.tops, vo))
    wp_p, dense_p, tops_p = _forward(model, dev_batch, precise=True)
    flips, flips_p = flips_vs_oracle(tops, orc_raw.tops, vo), flips_vs_oracle(tops_p, orc_raw.tops, vo)
    gap_p = first_flip_gap(tops_p, orc_raw.tops, vo)
    t_wp, _ = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    t_dev = t_wp.to(DEV)
    raw, raw_p, raw_f = (rel_err(x.cpu(), gold["waypoints"]) for x in (wp, wp_p, wp_f))
    gd = gold["waypoints"] - batch["gps"][:, -1:]
    raw_disp, raw_disp_p, raw_disp_f = (rel_err(x.cpu() - batch["gps"][:, -1:], gd) for x in (wp, wp_p, wp_f))
    dn = lambda x: rel_err(x.cpu(), gold["dense"]) if x is not None else float("nan")
    ade_f, fde_b_f, fde_1_f = R.ade(wp_f, t_dev).item(), R.fde(wp_f, t_dev).item(), R.fde(wp_f[-1:], t_dev[-1:]).item()
    n_problems = sum(int(t["top"].shape[0] * t["top"].shape[1]) for t in orc_raw.tops)
    log_parity(f"eval {name:30s} wp rel err vs reference golden: forced+precise {raw_f:.2e} | raw precise {raw_p:.2e} | raw tf32 {raw:.2e} || "
               f"displacements {raw_disp_f:.2e} / {raw_disp_p:.2e} / {raw_disp:.2e} || dense {dn(dense_f):.2e} / {dn(dense_p):.2e} / {dn(dense):.2e} || "
               f"flipped (b,h) problems of {n_problems}: precise {flips_p} (first-flip gap {gap_p:.1e}) tf32 {flips} || worst TF32 near-tie gap "
               f"{stats.get('worst_by_module')} ({stats['mismatches']} of {total} picks) || ADE gold {gold['ade']:.6f} forced+precise {ade_f:.6f} "
               f"raw tf32 {R.ade(wp, t_dev).item():.6f} | FDE gold {gold['fde_batch']:.6f} forced+precise {fde_b_f:.6f}")
    assert raw_f < 2e-5, raw_f
    assert abs(ade_f - gold["ade"]) < metric_tol(gold["ade"]), (ade_f, gold["ade"])
    assert abs(fde_b_f - gold["fde_batch"]) < metric_tol(gold["fde_batch"]), (fde_b_f, gold["fde_batch"])
    assert abs(fde_1_f - gold["fde"]) < metric_tol(gold["fde"]), (fde_1_f, gold["fde"])
    # A precise-mode flip must be a rounding-level tie.  With a visual stream the relevant rounding is fp16: the backbone plugin
    # returns its features in the input dtype (TimmBackbone.py:141-143), so a 1-ulp fp32 difference in the patch-embedding
    # accumulation can move a feature by half an fp16 ulp (2^-11 relative), which shifts the first frame-encoder measures by up
    # to ~1e-3 of max|M| (measured: 1.0e-3 over 111 616 problems); GPS-only models have no such stage (measured: no flip at all).
    assert gap_p <= (2e-3 if cfg.with_video else 1e-4), f"precise mode flipped a query that was not a rounding-level tie (gap {gap_p:.2e})"
    assert raw_p < (5e-3 if flips_p else 1e-3), (raw_p, flips_p)
    assert raw < (5e-3 if flips else 1e-3), (raw, flips)
    # the metric kernels themselves, on the product's own prediction
    assert abs(R.ade(wp, t_dev).item() - O.ade(wp.cpu(), t_wp).item()) < 5e-5 + 2e-7 * abs(gold["ade"])
    assert abs(R.fde(wp, t_dev).item() - O.fde(wp.cpu(), t_wp).item()) < 5e-5 + 2e-7 * abs(gold["fde_batch"])


# bf16 operand mode: the stated tolerance (north star: "bf16 operands with a stated looser tolerance").  bf16 keeps 8 mantissa bits
# (fp16, which the default path and the reference's autocast use for the patch embedding, keeps 11): the patch features move by
# ~2^-9 relative, and in inference they are additionally STORED in bf16 for the token convolution.  Waypoints: 4e-3 relative
# (BASELINE.md section 5 expects ~4e-3); displacement from the last observed position: 5e-2; ADE / FDE: 1e-2 relative.
BF16_WP_TOL, BF16_DISP_TOL, BF16_METRIC_TOL = 4e-3, 5e-2, 1e-2


@pytest.mark.parametrize("name", ["full_small_eval", "full_paper_eval", "dreyeve_paper_eval", "full_paper_b64_eval"])
def test_bf16_operand_mode(name):
    import routeformer_b200 as R
    from routeformer_b200 import ops

    gold = load_golden(name)
    cfg, spec, sd, batch = case_from_golden(gold)
    model = build_product(cfg, spec).to(DEV).eval()
    model.load_state_dict(sd)
    dev_batch = to_device(batch, DEV)
    wp32, _, tops32 = _forward(model, dev_batch, precise=False)
    launches = []
    orig = ops._gemm
    ops._gemm = lambda A, B, out, **k: (launches.append((A.dtype, out.dtype)), orig(A, B, out, **k))[1]
    try:
        with ops.bf16_operands():
            wp, _, tops = _forward(model, dev_batch, precise=False)
    finally:
        ops._gemm = orig
    # the patch embedding multiplies bf16 operands and stores bf16 features; the token convolution multiplies them as bf16
    assert (torch.bfloat16, torch.bfloat16) in launches and (torch.bfloat16, torch.float32) in launches
    assert not any(a == torch.float16 for a, _ in launches)
    last = batch["gps"][:, -1:]
    t_wp, _ = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    err = rel_err(wp.cpu(), gold["waypoints"])
    err_disp = rel_err(wp.cpu() - last, gold["waypoints"] - last)
    err_vs_tf32 = rel_err(wp.cpu(), wp32.cpu())
    ade, fde = R.ade(wp, t_wp.to(DEV)).item(), R.fde(wp, t_wp.to(DEV)).item()
    vo = view_order(cfg)
    flips = sum(int((a["top"].sort(-1).values != b["top"].sort(-1).values).any(-1).sum()) for a, b in zip(tops, tops32)
                if a["top"].shape == b["top"].shape)
    log_parity(f"bf16 {name:28s} wp rel err vs reference golden {err:.2e} (vs the TF32/fp16 default {err_vs_tf32:.2e}) | displacements {err_disp:.2e} | "
               f"ADE {ade:.6f} (gold {gold['ade']:.6f}) FDE {fde:.6f} (gold {gold['fde_batch']:.6f}) | (b,h) problems with a different top-u set "
               f"than the default mode: {flips}")
    assert err < BF16_WP_TOL, err
    assert err_disp < BF16_DISP_TOL, err_disp
    assert abs(ade - gold["ade"]) < BF16_METRIC_TOL * abs(gold["ade"]) and abs(fde - gold["fde_batch"]) < BF16_METRIC_TOL * abs(gold["fde_batch"])
    # training keeps fp32 features (the token convolution's weight gradient needs them): bf16 reaches the patch embedding only
    model.train()
    launches.clear()
    ops._gemm = lambda A, B, out, **k: (launches.append((A.dtype, out.dtype)), orig(A, B, out, **k))[1]
    try:
        with ops.bf16_operands():
            torch.manual_seed(1)
            out = model(dev_batch)
            (out[0] if isinstance(out, tuple) else out).sum().backward()
    finally:
        ops._gemm = orig
    assert (torch.bfloat16, torch.float32) in launches and (torch.bfloat16, torch.bfloat16) not in launches


def _train_step(sd, cfg, spec, batch, t_wp, t_dense):
    """One fwd+bwd on the GPU and on the oracle (replaying the GPU's top-u selections). Returns losses, grads, models."""
    import routeformer_b200 as R

    model = build_product(cfg, spec).to(DEV).train()
    model.load_state_dict(sd)
    model.record_tops = []
    lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
    torch.manual_seed(12345)
    wp, dense = model(to_device(batch, DEV))
    loss = lossf(wp, t_wp.to(DEV)) + 0.5 * lossf(dense, t_dense.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe") and
                                          not k.startswith("video_backbone")) for k, v in sd.items()}
    orc = O.Routeformer(params, cfg, spec)
    torch.manual_seed(12345)
    rwp, rdense = orc.forward(batch, training=True, draw=ReplayDraw(tops_for_oracle(model.record_tops, ["right", "left", "front"])))
    rloss = O.future_discounted_loss(rwp, t_wp) + 0.5 * O.future_discounted_loss(rdense, t_dense)
    rloss.backward()
    named = dict(model.named_parameters())
    rows = []
    for k, p in params.items():
        if p.requires_grad:
            g = named[k].grad
            assert g is not None and torch.isfinite(g).all(), f"missing / non-finite gradient for {k}"
            rows.append((k, p.grad.norm().item(), (g.cpu().double() - p.grad.double()).norm().item()))
    return model, orc, loss.item(), rloss.item(), rows


def test_train_step_gradients():
    """fwd+bwd of a training step: loss, every parameter gradient and the BatchNorm running statistics vs oracle autograd."""
    import os
    import statistics

    from tests.helpers import condition_weights

    gold = load_golden("full_small_train")
    cfg, spec, sd, batch = case_from_golden(gold)
    t_wp, t_dense = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    os.makedirs("gpurun_out", exist_ok=True)

    # (1) the golden problem (raw random weights): loss, BN statistics, all gradients present.  Gradients of this problem
    # are chaotic under TF32 (the reference itself moves by ~11% median, tools/tf32_sensitivity.py), so they are only logged.
    model, orc, loss, rloss, rows = _train_step(sd, cfg, spec, batch, t_wp, t_dense)
    assert abs(loss - rloss) < 2e-3 * abs(rloss)
    assert abs(loss - gold["loss"]) < 2e-2 * abs(gold["loss"])  # raw, vs the reference's own training step
    for k, v in orc.bn_updates.items():
        assert torch.allclose(model.state_dict()[k].cpu(), v, atol=1e-4, rtol=1e-3), k
    assert int(model.state_dict()["gps_backbone.encoder.conv_layers.0.norm.num_batches_tracked"]) == 1
    assert len(rows) > 200
    with open("gpurun_out/grad_table_golden.txt", "w") as f:
        for k, ref_n, err in rows:
            f.write(f"{k:90s} ref {ref_n:.4e} abs_err {err:.3e} rel {err / max(ref_n, 1e-30):.3e}\n")
    head = {k: err / ref_n for k, ref_n, err in rows if k.startswith("gps_backbone.decoder.projection")}
    assert max(head.values()) < 5e-3  # the last layer sees no upstream chaos

    # (2) conditioned weights: every gradient must match to TF32 accuracy (worst case = ReLU-mask flips, a few %)
    model, orc, loss, rloss, rows = _train_step(condition_weights(sd), cfg, spec, batch, t_wp, t_dense)
    assert abs(loss - rloss) < 1e-3 * abs(rloss)
    rel = []
    with open("gpurun_out/grad_table_conditioned.txt", "w") as f:
        for k, ref_n, err in rows:
            f.write(f"{k:90s} ref {ref_n:.4e} abs_err {err:.3e} rel {err / max(ref_n, 1e-30):.3e}\n")
            if ref_n > 1e-6:  # analytically-zero gradients (key biases, biases in front of BatchNorm) are pure noise
                rel.append((err / ref_n, k))
            else:
                assert err < 2e-3, k  # vs gradient norms of O(1..20) in the same layers
    rel.sort(reverse=True)
    assert rel[0][0] < 6e-2, rel[:5]
    assert statistics.median(r for r, _ in rel) < 1.5e-2


def _gpu_grads(sd, cfg, spec, batch, t_wp, t_dense, precise: bool, seed: int = 12345, forced=None):
    """One training-mode fwd+bwd on the GPU -> (model, loss, {name: grad clone}, recorded selections)."""
    import routeformer_b200 as R
    from routeformer_b200 import ops

    model = build_product(cfg, spec).to(DEV).train()
    model.load_state_dict(sd)
    model.record_tops = []
    model.forced_tops = forced
    lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
    torch.manual_seed(seed)
    with ops.precise(precise):
        wp, dense = model(to_device(batch, DEV))
        loss = lossf(wp, t_wp.to(DEV)) + 0.5 * lossf(dense, t_dense.to(DEV))
        loss.backward()
    torch.cuda.synchronize()
    model.forced_tops = None
    grads = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
    return model, loss.item(), grads, model.record_tops


def _grad_errors(gold, grads, zero_tol=1e-4):
    """(per-parameter gradient-norm relative errors, full-tensor relative errors) against the reference's training step.
    zero_tol: absolute bound on the analytically-zero gradients -- what is left there is the rounding residue of sums whose terms
    cancel, so it scales with the operand rounding (TF32: 2^-11 per product; measured 1.2e-4 on norm2.bias, whose companion
    weight gradient has entries of 0.1-1; precise mode: fp32 level)."""
    rel_n, full = [], []
    for k, n in gold["grad_norm"].items():
        if k.startswith("video_backbone"):
            continue  # frozen in the product (reference: trained only after epoch 10, TimmBackbone.py:123)
        assert k in grads, f"missing gradient for {k}"
        if n > 1e-5:
            rel_n.append((abs(grads[k].norm().item() - n) / n, k))
    for k, g in gold["grad_small"].items():
        if k.startswith("video_backbone"):
            continue
        if g.norm() > 1e-5:
            full.append((rel_err(grads[k], g), k))
        else:  # analytically-zero gradients (key biases: softmax shift invariance; biases in front of BatchNorm)
            assert (grads[k] - g).abs().max() < zero_tol, k
    rel_n.sort(reverse=True)
    full.sort(reverse=True)
    return rel_n, full


@pytest.mark.parametrize("name", ["full_small_train", "full_paper_train"])
def test_train_step_raw_against_reference(name):
    """Gradient parity at the reference's OWN initialisation (no conditioning of the weights) against the golden of the
    unmodified reference's training step: loss, every per-parameter gradient norm, the stored gradient tensors, the BatchNorm
    statistics.  `full_paper_train` is the paper configuration at B = 8 (BASELINE configs[2] shard shape).
      (i)   precise mode (3xTF32) with the reference's top-u selections forced: arithmetic parity, asserted tightly;
      (ii)  precise mode with the kernels' own selections (raw): fp32-level ties flip a few queries (first-flip gap logged);
      (iii) default TF32 mode (raw): the reference's own gradients move by ~11 % (median) under TF32 operand rounding at this
            initialisation (tools/tf32_sensitivity.py) -- logged, loss asserted."""
    import statistics

    gold = load_golden(name)
    cfg, spec, sd, batch = case_from_golden(gold)
    t_wp, t_dense = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    params = {k: v.clone() for k, v in sd.items()}
    orc_raw = O.Routeformer(params, cfg, spec)
    torch.manual_seed(12345)
    with torch.no_grad():
        orc_raw.forward(batch, training=True)
    vo = ["right", "left", "front"]
    model, loss_f, grads_f, _ = _gpu_grads(sd, cfg, spec, batch, t_wp, t_dense, precise=True, forced=tops_for_product(orc_raw.tops, vo))
    _, loss_p, grads_p, tops_p = _gpu_grads(sd, cfg, spec, batch, t_wp, t_dense, precise=True)
    _, loss_t, grads_t, tops_t = _gpu_grads(sd, cfg, spec, batch, t_wp, t_dense, precise=False)
    med = lambda rows: statistics.median(r for r, _ in rows)
    rows = {}
    for tag, grads in (("forced+precise", grads_f), ("raw precise", grads_p), ("raw tf32", grads_t)):
        rel_n, full = _grad_errors(gold, grads, zero_tol=3e-4 if tag == "raw tf32" else 1e-4)
        rows[tag] = (rel_n, full)
        by_mod = {}
        for r, k in full:
            by_mod.setdefault(k.split(".")[0], []).append(r)
        by_mod = {m: f"{statistics.median(v):.1e}" for m, v in by_mod.items()}
        log_parity(f"train {name:18s} {tag:15s} grad-norm rel err: median {med(rel_n):.2e} max {rel_n[0][0]:.2e} ({rel_n[0][1]}) | full tensors: "
                   f"median {med(full):.2e} max {full[0][0]:.2e} ({full[0][1]}) | full-tensor median by module {by_mod}")
        if tag == "forced+precise":
            for r, k in full[:6]:
                log_parity(f"      worst {k}: rel {r:.2e} |ref| {gold['grad_small'][k].norm():.3e} |err| {(grads[k] - gold['grad_small'][k]).norm():.3e}")
    log_parity(f"train {name:18s} loss gold {gold['loss']:.6f} forced+precise {loss_f:.6f} raw precise {loss_p:.6f} raw tf32 {loss_t:.6f} | flipped "
               f"(b,h) problems: precise {flips_vs_oracle(tops_p, orc_raw.tops, vo)} (first-flip gap {first_flip_gap(tops_p, orc_raw.tops, vo):.1e}) "
               f"tf32 {flips_vs_oracle(tops_t, orc_raw.tops, vo)}")
    assert abs(loss_f - gold["loss"]) < 2e-5 * abs(gold["loss"])
    assert abs(loss_p - gold["loss"]) < 1e-3 * abs(gold["loss"])
    assert abs(loss_t - gold["loss"]) < 2e-2 * abs(gold["loss"])
    rel_n, full = rows["forced+precise"]
    assert med(rel_n) < 5e-3 and rel_n[0][0] < 5e-2, rel_n[:5]
    assert med(full) < 1e-2 and full[0][0] < 1e-1, full[:5]
    assert first_flip_gap(tops_p, orc_raw.tops, vo) <= 2e-3
    new_sd = model.state_dict()
    for k, v in gold["bn"].items():
        if "num_batches" in k:
            assert int(new_sd[k]) == int(v), k
        else:
            assert torch.allclose(new_sd[k].cpu(), v, atol=1e-5, rtol=1e-4), k


def test_train_backbone_gradients():
    """train_backbone (the reference un-freezes its visual backbone after epoch 10, TimmBackbone.py:123): the patch-embedding
    projection receives its weight / bias gradient through the wgrad GEMM.  Against the reference's own training-step golden
    (whose backbone stub is trainable), precise mode with the reference's selections forced."""
    import routeformer_b200 as R
    from routeformer_b200 import ops

    gold = load_golden("full_small_train")
    cfg, spec, sd, batch = case_from_golden(gold)
    t_wp, t_dense = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    orc_raw = O.Routeformer({k: v.clone() for k, v in sd.items()}, cfg, spec)
    torch.manual_seed(12345)
    with torch.no_grad():
        orc_raw.forward(batch, training=True)
    model = build_product(cfg, spec)
    model.video_backbone.proj.requires_grad_(True)
    model = model.to(DEV).train()
    model.load_state_dict(sd)
    model.forced_tops = tops_for_product(orc_raw.tops, ["right", "left", "front"])
    lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
    torch.manual_seed(12345)
    with ops.precise(True):
        wp, dense = model(to_device(batch, DEV))
        loss = lossf(wp, t_wp.to(DEV)) + 0.5 * lossf(dense, t_dense.to(DEV))
        loss.backward()
    torch.cuda.synchronize()
    model.forced_tops = None
    # with a trainable backbone the fp16 clips take the TF32 patch GEMM instead of the fp16 one: same loss to TF32 accuracy
    assert abs(loss.item() - gold["loss"]) < 1e-3 * abs(gold["loss"])
    gw, gb = model.video_backbone.proj.weight.grad, model.video_backbone.proj.bias.grad
    assert gw is not None and gb is not None
    nw, nb = gold["grad_norm"]["video_backbone.proj.weight"], gold["grad_norm"]["video_backbone.proj.bias"]
    e_w, e_b = abs(gw.norm().item() - nw) / nw, abs(gb.norm().item() - nb) / nb
    e_full = rel_err(gb.cpu(), gold["grad_small"]["video_backbone.proj.bias"])
    log_parity(f"train_backbone: patch-embedding gradient norms vs reference: weight rel err {e_w:.2e}, bias {e_b:.2e}; bias tensor rel err {e_full:.2e}")
    assert e_w < 2e-2 and e_b < 2e-2 and e_full < 5e-2
    # every other gradient is still there
    assert all(p.grad is not None for n, p in model.named_parameters() if p.requires_grad)


def test_two_shard_data_parallel_equivalence():
    """the global batch split over two ranks (contiguous shards, identical weights, identical CPU draws, per-shard
    BatchNorm statistics), gradients summed and scaled by 1/world == the oracle's per-shard autograd averaged the same way.
    Both "ranks" run through the real model on one GPU; the collective itself is covered by the gloo tests."""
    from routeformer_b200.parallel import shard_batch
    from tests.helpers import condition_weights

    gold = load_golden("full_small_train")
    cfg, spec, sd, _ = case_from_golden(gold)
    sd = condition_weights(sd)
    B, world = 4, 2
    batch = O.synthetic_batch(B, cfg, gold["shapes"], seed=31)
    t_wp, t_dense = targets_for(cfg, B, 32)
    gpu_sum, ref_sum, losses = {}, {}, []
    for rank in range(world):
        shard = shard_batch(batch, rank, world)
        tw, td = t_wp[rank * 2:(rank + 1) * 2], t_dense[rank * 2:(rank + 1) * 2]
        model, loss, grads, tops = _gpu_grads(sd, cfg, spec, shard, tw, td, precise=False)
        params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe") and
                                              not k.startswith("video_backbone")) for k, v in sd.items()}
        orc = O.Routeformer(params, cfg, spec)
        torch.manual_seed(12345)
        rwp, rdense = orc.forward(shard, training=True, draw=ReplayDraw(tops_for_oracle(tops, ["right", "left", "front"])))
        rloss = O.future_discounted_loss(rwp, tw) + 0.5 * O.future_discounted_loss(rdense, td)
        rloss.backward()
        losses.append((loss, rloss.item()))
        for k, p in params.items():
            if p.requires_grad:
                gpu_sum[k] = gpu_sum.get(k, 0) + grads[k] / world
                ref_sum[k] = ref_sum.get(k, 0) + p.grad / world
    for l, r in losses:
        assert abs(l - r) < 1e-3 * abs(r)
    rel = sorted(((rel_err(gpu_sum[k], ref_sum[k]), k) for k in ref_sum if ref_sum[k].norm() > 1e-6), reverse=True)
    import statistics
    log_parity(f"dp2   two-shard averaged gradients vs oracle: median rel err {statistics.median(r for r, _ in rel):.2e} max {rel[0][0]:.2e} ({rel[0][1]})")
    # (the worst parameters are the distil convolution's bias / weight: their gradients pass through batch statistics of TWO rows
    # per shard, the most TF32-sensitive sums of the model; measured 5.6e-2 .. 6.0e-2 depending on the crop kernel's summation order)
    assert rel[0][0] < 8e-2 and statistics.median(r for r, _ in rel) < 1.5e-2, rel[:5]


def test_standalone_perceive_modules():
    """`PerceiveEncoder(...)(x)` / `PerceiveDecoder(...)(x_enc, x_dec)` used stand-alone, as the baselines of
    experiments/gimo/adapted_gimo.py:59-67 and experiments/multimodal_transformer do -- same class names and constructor
    arguments, the module draws its own index tables on the CPU generator (LiveIndexSource).  Against the reference's outputs."""
    import routeformer_b200 as R
    from routeformer_b200 import ops

    gold = load_golden("submodules")
    enc = R.PerceiveEncoder(in_channels=40, out_len=1, out_channels=24, n_heads=8, layers=2, d_ff=64, dropout=0.0).to(DEV).eval()
    tmpl = {}
    O._perceive_encoder_keys(tmpl, "e", 40, 24, 128, 2, 64)
    sd = O.fill_state_dict({k[2:]: v for k, v in tmpl.items()}, 21)
    assert set(sd) == set(enc.state_dict())
    enc.load_state_dict(sd)
    dec = R.PerceiveDecoder(query_channels=16, value_channels=16, out_channels=16, out_len=12, dropout=0.0, d_ff=32, n_heads=4, layers=2,
                            mix=False).to(DEV).eval()
    tmpl = {}
    O._perceive_decoder_keys(tmpl, "d", 16, 16, 16, 2, 32)
    sd = O.fill_state_dict({k[2:]: v for k, v in tmpl.items()}, 22)
    assert set(sd) == set(dec.state_dict())
    dec.load_state_dict(sd)
    ge, gd = gold["perceive_encoder"], gold["perceive_decoder"]
    for precise, tol in ((True, 1e-4), (False, 5e-3)):
        with torch.no_grad(), ops.precise(precise):
            torch.manual_seed(3)
            y = enc(ge["x"].to(DEV))
            torch.manual_seed(4)
            z = dec(gd["x_enc"].to(DEV), gd["x_dec"].to(DEV))
        torch.cuda.synchronize()
        assert y.shape == ge["y"].shape and z.shape == gd["y"].shape
        e_enc, e_dec = rel_err(y.cpu(), ge["y"]), rel_err(z.cpu(), gd["y"])
        log_parity(f"perceive stand-alone ({'precise' if precise else 'tf32'}): encoder rel err {e_enc:.2e}, decoder rel err {e_dec:.2e}")
        assert e_enc < tol and e_dec < tol, (precise, e_enc, e_dec)


def test_gaze_centred_fov_and_plugin_api():
    """North-star extension: gaze-centred FoV window for the front camera; and the VideoBackboneModule plugin contract."""
    gold = load_golden("full_small_eval")
    cfg, spec, sd, batch = case_from_golden(gold)
    model = build_product(cfg, spec, fov="gaze").to(DEV).eval()
    model.load_state_dict(sd)
    model.record_tops = []
    torch.manual_seed(7)
    with torch.no_grad():
        wp, dense = model(to_device(batch, DEV))
    orc = O.Routeformer(sd, cfg, spec, fov="gaze")
    Draw = ReplayDraw(tops_for_oracle(model.record_tops, ["right", "left", "front"]))

    torch.manual_seed(7)
    with torch.no_grad():
        rwp, rdense = orc.forward(batch, draw=Draw)
    assert rel_err(wp.cpu(), rwp) < 1e-3 and rel_err(dense.cpu(), rdense) < 1e-2
    # plugin API: [N,3,H,W] -> [N,C,G,G] in the input dtype
    frames = batch["front_video"][0, :5].to(DEV)
    feats = model.video_backbone(frames)
    ref = O.patch_backbone(sd, "video_backbone", batch["front_video"][0, :5], spec).to(frames.dtype)
    assert feats.shape == ref.shape and feats.dtype == frames.dtype
    assert rel_err(feats.float().cpu(), ref.float()) < 3e-3


def test_uint8_frames_equal_the_loaders_fp16_frames():
    """SURVEY 8(f) N4 at model level (`bench.py --u8-frames`): raw uint8 frames, staged by `stage_batch` (only the consumed
    frames travel) and converted in the crop kernel, give the SAME BITS as the fp16 frames the reference's loader makes of them
    (`astype(float16) / 255`, io/dataset.py:1522).  The frames only enter the crop kernel (frozen backbone), so identical patch
    bits make everything downstream -- the training step included -- the same computation."""
    gold = load_golden("full_small_eval")
    cfg, spec, sd, batch = case_from_golden(gold)
    raw = {k: ((v.float() * 255.0).round().clamp(0, 255).to(torch.uint8) if k.endswith("_video") else v) for k, v in batch.items()}
    loader = {k: (torch.from_numpy(v.numpy().astype("float16") / 255.0) if k.endswith("_video") else v) for k, v in raw.items()}
    assert all(loader[k].dtype == torch.float16 for k in loader if k.endswith("_video"))
    model = build_product(cfg, spec, fov="gaze").to(DEV).eval()
    model.load_state_dict(sd)
    outs = []
    for host in (raw, loader):
        staged = model.stage_batch({k: v.contiguous().pin_memory() for k, v in host.items()}, torch.device(DEV))
        torch.manual_seed(3)
        with torch.no_grad():
            outs.append(model(staged))
    assert outs[0][0].dtype == torch.float32
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # the unstaged uint8 clip (all 40 frames resident) reads the same frames
    torch.manual_seed(3)
    with torch.no_grad():
        wp, dense = model(to_device(raw, DEV))
    assert torch.equal(wp, outs[0][0]) and torch.equal(dense, outs[0][1])


def test_target_pass_and_errors():
    """preprocess_batch(target, training=False) on the 30 target frames (full_comparison.py:482) + error behaviour."""
    gold = load_golden("full_small_eval")
    cfg, spec, sd, _ = case_from_golden(gold)
    model = build_product(cfg, spec).to(DEV).eval()
    model.load_state_dict(sd)
    tgt = O.synthetic_batch(2, cfg, "tiny", seed=8, T=30)
    model.record_tops = []
    torch.manual_seed(1)
    with torch.no_grad():
        motion, visual = model.preprocess_batch(to_device(tgt, DEV), training=False)
    # 4 streams x 30 frames = 120 tokens, the encoder returns the last seq_len = 40 (the caller slices [:, :30])
    assert motion.shape == (2, 30, 2) and visual.shape == (2, 40, cfg.encoder_hidden_size)
    Draw = ReplayDraw(tops_for_oracle(model.record_tops, ["right", "left", "front"]))

    torch.manual_seed(1)
    with torch.no_grad():
        rm, rv = O.Routeformer(sd, cfg, spec).preprocess(tgt, False, Draw)
    assert rel_err(motion.cpu(), rm) < 1e-6 and rel_err(visual.cpu(), rv) < 1e-2
    with pytest.raises(RuntimeError):  # no CPU fallback
        model(tgt)
    short = {**to_device(tgt, DEV), "gaze": torch.rand(2, 40, 2, device=DEV)}
    with pytest.raises(Exception):  # utils/filter.py:20-21: target length must be < samples
        model.preprocess_batch(short, training=False)


def test_cuda_graph_step_matches_eager():
    """Replaying the captured graph of zero-grad + forward + loss + backward gives the same loss and gradients as the eager
    path for consecutive steps (the CPU draws of every step reach the graph through the persistent pinned buffer).
    Compared before the optimiser: Adam turns noise-level gradients (e.g. the analytically-zero key biases) into +-lr updates."""
    import routeformer_b200 as R
    from routeformer_b200.parallel import DataParallelTrainer

    gold = load_golden("full_small_train")
    cfg, spec, sd, batch = case_from_golden(gold)
    t_wp, t_dense = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
    results = []
    for graph in (False, False, True):
        model = build_product(cfg, spec).to(DEV).train()
        model.load_state_dict(sd)
        trainer = DataParallelTrainer(model, lambda out, tgt: lossf(out[0], tgt[0]) + 0.5 * lossf(out[1], tgt[1]), use_cuda_graph=graph)
        dev_batch = to_device(batch, DEV)
        tgt = (t_wp.to(DEV), t_dense.to(DEV))
        if graph:
            trainer._capture(dev_batch, tgt)  # consumes CPU draws (2 warm-up passes + the capture pass)
        torch.manual_seed(99)
        steps = []
        for _ in range(3):
            loss = trainer._replay(dev_batch, tgt) if graph else trainer._fwd_bwd(dev_batch, tgt)
            torch.cuda.synchronize()
            steps.append((loss.item(), trainer.arena.grad.clone(), list(model.last_draw_log)))
        results.append(steps)
    eager, eager2, graphed = results
    for (l0, g0, d0), (l2, g2, _), (l1, g1, d1) in zip(eager, eager2, graphed):
        assert d0 == d1
        assert abs(l0 - l1) < 1e-5 * abs(l0), (l0, l1)
        # wgrad / bias-grad reductions use fp32 atomics (scheduling-order dependent): the graph must agree with the eager path
        # as well as two eager runs agree with each other
        run_to_run = rel_err(g2.cpu(), g0.cpu())
        assert rel_err(g1.cpu(), g0.cpu()) <= max(5 * run_to_run, 1e-5), (rel_err(g1.cpu(), g0.cpu()), run_to_run)
    assert eager[0][0] != eager[1][0]  # different draws every step


def test_cuda_graph_with_paper_dropouts():
    """The paper configuration trains with view / gaze / feature dropout (full_comparison.py:272-275).  Under CUDA graphs:
    one graph per (left, right, gaze) drop pattern, chosen per step by the same CPU torch.rand draws the reference makes;
    feature-dropout masks come from a device-resident step counter the graph itself advances.  The graph path must follow the
    eager path step for step: same patterns, same draws, same losses, same gradients."""
    import dataclasses

    import routeformer_b200 as R
    from routeformer_b200.parallel import DataParallelTrainer

    gold = load_golden("full_small_train")
    cfg, spec, sd, batch = case_from_golden(gold)
    cfg = dataclasses.replace(cfg, view_dropout=0.6, gaze_dropout=0.4)
    t_wp, t_dense = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
    runs = []
    for graph in (False, True):
        from routeformer_b200 import ops
        ops.DropoutStream._base.clear()
        ops.DropoutStream._base_host.clear()
        model = build_product(cfg, spec, feature_dropout=0.1).to(DEV).train()
        model.load_state_dict(sd)
        trainer = DataParallelTrainer(model, lambda out, tgt: lossf(out[0], tgt[0]) + 0.5 * lossf(out[1], tgt[1]), use_cuda_graph=graph)
        dev_batch, tgt = to_device(batch, DEV), (t_wp.to(DEV), t_dense.to(DEV))
        torch.manual_seed(4242)
        steps = []
        for _ in range(10):
            loss = trainer._replay(dev_batch, tgt) if graph else trainer._fwd_bwd(dev_batch, tgt)
            torch.cuda.synchronize()
            steps.append((model.last_pattern if not graph else trainer._next_pattern, loss.item(), trainer.arena.grad.clone(),
                          list(model.last_draw_log)))
        runs.append((steps, trainer))
    (eager, _), (graphed, trainer) = runs
    assert len(trainer._graphs) > 1, "the seeds must exercise more than one drop pattern"
    assert len({p for p, *_ in eager}) == len(trainer._graphs)
    for (p0, l0, g0, d0), (p1, l1, g1, d1) in zip(eager, graphed):
        assert p0 == p1 and d0 == d1
        assert abs(l0 - l1) < 1e-5 * abs(l0), (l0, l1)
        assert rel_err(g1.cpu(), g0.cpu()) < 1e-3
    assert len({round(l, 6) for _, l, _, _ in graphed}) == len(graphed)  # fresh masks / draws every replay


def test_training_step_under_cuda_graph():
    """SURVEY 8(f) N1 at speed: the reference's whole training_step (forward + eval-mode target pass + both losses + detached
    dense re-weighting, full_comparison.py:470-532) captured as one graph through DataParallelTrainer(step_fn, draw_fn)."""
    import routeformer_b200 as R
    from routeformer_b200.parallel import DataParallelTrainer

    results = []
    for graph in (False, True):
        gold, cfg, spec, sd, batch, model, dev_batch = _steps_case()
        model.train()
        steps = R.ParallelTrainerSteps(model)
        step_fn, draw_fn = steps.graph_hooks(current_epoch=10)
        trainer = DataParallelTrainer(model, None, use_cuda_graph=graph, step_fn=step_fn, draw_fn=draw_fn)
        trainer.graph_tag = True
        torch.manual_seed(12345)
        out = []
        for _ in range(3):
            loss = trainer._replay(dev_batch, ()) if graph else trainer._fwd_bwd(dev_batch, ())
            torch.cuda.synchronize()
            out.append((loss.item(), steps.last_metrics["train_ade"].item(), trainer.arena.grad.clone()))
        results.append(out)
    g = gold["train_epoch10"]
    assert abs(results[0][0][0] - g["loss"]) < 2e-2 * abs(g["loss"])  # first eager step = the reference's golden step
    for (l0, a0, g0), (l1, a1, g1) in zip(*results):
        assert abs(l0 - l1) < 1e-5 * abs(l0) and abs(a0 - a1) < 1e-5 * abs(a0)
        assert rel_err(g1.cpu(), g0.cpu()) < 1e-3


def _steps_case():
    gold = load_golden("steps_small")
    cfg, spec = O.OracleConfig(**gold["cfg"]), O.BackboneSpec(**gold["spec"])
    sd = O.fill_state_dict(O.state_dict_template(cfg, spec), gold["wseed"])
    batch = {"train": O.synthetic_batch(gold["B"], cfg, "tiny", seed=gold["dseed"]),
             "target": O.synthetic_batch(gold["B"], cfg, "tiny", seed=gold["dseed"] + 1, T=cfg.pred_len)}
    batch["target"]["gps"] = batch["target"]["gps"] + batch["train"]["gps"][:, -1:]
    model = build_product(cfg, spec, discount_factor=gold["discount"], epsilon=gold["epsilon"], visual_epsilon=gold["visual_epsilon"],
                          dense_loss_ratio=gold["dense_loss_ratio"]).to(DEV)
    model.load_state_dict(sd)
    dev_batch = {k: to_device(v, DEV) for k, v in batch.items()}
    return gold, cfg, spec, sd, batch, model, dev_batch


@pytest.mark.parametrize("epoch", [0, 10])
def test_training_step_caller(epoch):
    """ParallelTrainerSteps.training_step (full_comparison.py:470-532) vs the oracle (replaying the GPU's selections) and the
    reference's golden values: forward, eval-mode target pass, both losses, the detached dense re-weighting, ADE / FDE."""
    import routeformer_b200 as R

    gold, cfg, spec, sd, batch, model, dev_batch = _steps_case()
    model.train()
    model.record_tops = []
    steps = R.ParallelTrainerSteps(model)
    torch.manual_seed(12345)
    loss, metrics = steps.training_step(dev_batch, current_epoch=epoch)
    loss.backward()
    torch.cuda.synchronize()
    g = gold[f"train_epoch{epoch}"]
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe") and
                                          not k.startswith("video_backbone")) for k, v in sd.items()}
    orc = O.Routeformer(params, cfg, spec)
    torch.manual_seed(12345)
    rloss, rm = O.training_step(orc, batch, gold["discount"][0], gold["dense_loss_ratio"], epoch,
                                ReplayDraw(tops_for_oracle(model.record_tops, ["right", "left", "front"])))
    assert abs(loss.item() - rloss.item()) < 2e-3 * abs(rloss.item())
    assert abs(metrics["train_loss"].item() - rm["trajectory_loss"].item()) < 2e-3 * abs(rm["trajectory_loss"].item())
    assert abs(metrics["train_dense_loss"].item() - rm["dense_loss"].item()) < 1e-2 * abs(rm["dense_loss"].item())
    assert abs(metrics["train_ade"].item() - rm["ade"].item()) < 2e-3 * rm["ade"].item()
    assert abs(metrics["train_fde"].item() - rm["fde"].item()) < 2e-3 * rm["fde"].item()
    # raw, against the reference's own numbers (no replay)
    assert abs(loss.item() - g["loss"]) < 2e-2 * abs(g["loss"])
    assert abs(metrics["train_ade"].item() - g["ade"]) < 1e-2 * g["ade"]
    named = dict(model.named_parameters())
    missing = [k for k, p in params.items() if p.requires_grad and (named[k].grad is None or not torch.isfinite(named[k].grad).all())]
    assert not missing, missing


@pytest.mark.parametrize("batched", [True, False])
def test_eval_step_caller(batched):
    """ParallelTrainerSteps.eval_step (full_comparison.py:654-679): RNG order over the five forwards, mean prediction, per-clip
    metrics -- as ONE five-fold batch (`Routeformer.forward_samples`, the default) and as the reference's loop of five forwards."""
    import routeformer_b200 as R

    gold, cfg, spec, sd, batch, model, dev_batch = _steps_case()
    model.eval()
    steps = R.ParallelTrainerSteps(model)
    steps.batched_samples = batched
    logs = []
    orig = model.prepare_draws

    def spy(*a, **k):
        plan = orig(*a, **k)
        logs.extend(model.last_draw_log)
        return plan

    model.prepare_draws = spy
    losses, ades, fdes = steps.eval_step(dev_batch)
    torch.cuda.synchronize()
    if batched:
        assert not logs  # one plan for all five forwards
        logs = model.last_draw_log
    e = gold["eval"]
    assert [(lk, (lq, u)) for lk, lq, u in logs] == [tuple(d) for d in e["draws"]]
    assert rel_err(steps.last_mean_prediction.cpu(), e["mean_prediction"]) < 5e-3
    assert torch.allclose(ades.cpu(), e["ades"], rtol=1e-2) and torch.allclose(fdes.cpu(), e["fdes"], rtol=1e-2)
    assert torch.allclose(losses.cpu(), e["losses"], rtol=1e-2)
    # the metric kernel itself, exactly, on the product's own mean prediction
    mean = steps.last_mean_prediction.cpu()
    t = batch["target"]["gps"]
    ref_ade = torch.stack([O.ade(mean[i:i + 1], t[i:i + 1]) for i in range(mean.shape[0])])
    assert torch.allclose(ades.cpu(), ref_ade, rtol=2e-5)


@pytest.mark.parametrize("case", ["full_small_eval", "dreyeve_small", "no_scene_small", "gps_only_paper", "full_paper_b64_eval"])
def test_forward_samples_equals_sequential_forwards(case):
    """`forward_samples(batch, 5)` against five sequential eval forwards after the same `torch.manual_seed`: identical CPU draw
    log, every sample's waypoints / dense features equal to its sequential forward (the arithmetic per sequence is the same;
    only the batch the kernels see is five times larger)."""
    gold = load_golden(case)
    cfg, spec, sd, batch = case_from_golden(gold)
    model = build_product(cfg, spec).to(DEV).eval()
    model.load_state_dict(sd)
    dev_batch = {k: v.to(DEV) for k, v in batch.items()}
    n = 5
    torch.manual_seed(12345)
    seq, seq_log = [], []
    with torch.no_grad():
        for _ in range(n):
            seq.append(model(dev_batch))
            seq_log += model.last_draw_log
    torch.manual_seed(12345)
    wp, dense = model.forward_samples(dev_batch, n)
    torch.cuda.synchronize()
    assert model.last_draw_log == seq_log
    ref_wp = torch.stack([(o[0] if cfg.dense_prediction else o) for o in seq]).cpu()   # [n, B, P, 2]
    got = wp.cpu()
    per_clip = ((got - ref_wp).flatten(2).norm(dim=2) / ref_wp.flatten(2).norm(dim=2).clamp_min(1e-12)).flatten()
    if bool(cfg.dense_prediction) and model.with_video:
        ref_d = torch.stack([o[1] for o in seq]).cpu()
        d_err = ((dense.cpu() - ref_d).flatten(2).norm(dim=2) / ref_d.flatten(2).norm(dim=2).clamp_min(1e-12)).flatten()
        per_clip = torch.maximum(per_clip, d_err)
    moved = int((per_clip > 1e-4).sum())
    print(f"forward_samples vs sequential [{case}]: median {float(per_clip.median()):.2e}, max {float(per_clip.max()):.2e}, "
          f"{moved} of {per_clip.numel()} (sample, clip) pairs above 1e-4")
    # Same arithmetic per sequence: the kernel variants the launch heuristics pick for a five-fold batch (thread count of the
    # generic attention kernel, tile schedule of the GEMMs) keep the summation order, so the samples agree bit for bit -- which
    # matters because ProbSparse's top-u selection would turn a rounding-level difference into a different selected query.
    assert moved == 0 and float(per_clip.max()) < 1e-6, (moved, float(per_clip.max()))
    assert wp.shape[:2] == (n, batch["gps"].shape[0])


def test_training_step_with_feature_dropout():
    """Training-mode feature dropout of the Perceive modules (paper config: 0.05; here 0.2 so that it matters): the oracle replays
    the CUDA path's Philox masks and top-u selections, so loss and predictions must agree at TF32 level; eval mode is unaffected."""
    import routeformer_b200 as R
    from routeformer_b200 import ops

    gold = load_golden("full_small_train")
    cfg, spec, sd, batch = case_from_golden(gold)
    t_wp, t_dense = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    p = 0.2
    model = build_product(cfg, spec, feature_dropout=p).to(DEV).train()
    model.load_state_dict(sd)
    model.record_tops = []
    ops.DropoutStream.log = []
    try:
        lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
        torch.manual_seed(12345)
        wp, dense = model(to_device(batch, DEV))
        loss = lossf(wp, t_wp.to(DEV)) + 0.5 * lossf(dense, t_dense.to(DEV))
        loss.backward()
        torch.cuda.synchronize()
        log = list(ops.DropoutStream.log)
    finally:
        ops.DropoutStream.log = None
    # 2 frame-encoder + gaze-encoder + video-encoder layers x 3 sites, 2 decoder layers x 5 sites
    assert len(log) == 3 * (3 * cfg.encoder_layers) + 5 * cfg.cross_modal_decoder_layers
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe") and
                                          not k.startswith("video_backbone")) for k, v in sd.items()}
    orc = O.Routeformer(params, cfg, spec)
    hook = ReplayDropout(log, p)
    torch.manual_seed(12345)
    with O.dropout_hook(hook):
        rwp, rdense = orc.forward(batch, training=True, draw=ReplayDraw(tops_for_oracle(model.record_tops, ["right", "left", "front"])))
    assert hook.exhausted()
    rloss = O.future_discounted_loss(rwp, t_wp) + 0.5 * O.future_discounted_loss(rdense, t_dense)
    assert rel_err(wp.detach().cpu(), rwp.detach()) < 1e-3
    assert rel_err(dense.detach().cpu(), rdense.detach()) < 2e-2
    assert abs(loss.item() - rloss.item()) < 2e-3 * abs(rloss.item())
    # dropout changes the result (vs the golden, dropout-free loss) and every trainable parameter still gets a finite gradient
    assert abs(loss.item() - gold["loss"]) > 1e-4 * abs(gold["loss"])
    named = dict(model.named_parameters())
    assert all(named[k].grad is not None and torch.isfinite(named[k].grad).all() for k, v in params.items() if v.requires_grad)
    # gradients against the oracle's autograd through the same masks (norm-wise; TF32 + top-u chaos bound as in the raw-weight test)
    rloss.backward()
    late = [k for k, v in params.items() if v.requires_grad and ("gps_backbone.projection" in k or "video_encoder.projection" in k)]
    assert late
    worst = max(((named[k].grad.cpu() - params[k].grad).norm() / params[k].grad.norm().clamp_min(1e-6)).item() for k in late)
    assert worst < 0.1, worst
    # eval mode: no dropout, identical to the dropout-free model (state reloaded: the training pass moved the BatchNorm statistics)
    model.load_state_dict(sd)
    model.eval()
    ref_model = build_product(cfg, spec).to(DEV).eval()
    ref_model.load_state_dict(sd)
    with torch.no_grad():
        torch.manual_seed(1)
        a = model(to_device(batch, DEV))[0]
        torch.manual_seed(1)
        b = ref_model(to_device(batch, DEV))[0]
    assert torch.equal(a, b)


def test_informer_dropout():
    """GPSBackboneConfig.dropout > 0 (the paper configuration uses 0.0, full_comparison.py:170): DataEmbedding dropout and the
    three dropout sites of every Informer encoder / decoder layer, with the CUDA masks replayed inside the oracle."""
    import routeformer_b200 as R
    from routeformer_b200 import ops

    gold = load_golden("full_small_train")
    cfg, spec, sd, batch = case_from_golden(gold)
    t_wp, t_dense = targets_for(cfg, gold["B"], gold["dseed"] + 1000)
    p = 0.15
    model = build_product(cfg, spec, informer_dropout=p).to(DEV).train()
    model.load_state_dict(sd)
    model.record_tops = []
    ops.DropoutStream.log = []
    try:
        lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
        torch.manual_seed(12345)
        wp, dense = model(to_device(batch, DEV))
        loss = lossf(wp, t_wp.to(DEV)) + 0.5 * lossf(dense, t_dense.to(DEV))
        loss.backward()
        torch.cuda.synchronize()
        log = list(ops.DropoutStream.log)
    finally:
        ops.DropoutStream.log = None
    # 2 embeddings + 3 sites per encoder layer + 5 per decoder layer (self out, cross out, ffn hidden, ffn out; no prob dropout: ProbAttention)
    assert len(log) == 2 + 3 * cfg.e_layers + 4 * cfg.d_layers, len(log)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe") and
                                          not k.startswith("video_backbone")) for k, v in sd.items()}
    orc = O.Routeformer(params, cfg, spec)
    hook = ReplayDropout(log, p)
    torch.manual_seed(12345)
    with O.dropout_hook(hook, informer=True, perceive=False):
        rwp, rdense = orc.forward(batch, training=True, draw=ReplayDraw(tops_for_oracle(model.record_tops, ["right", "left", "front"])))
    assert hook.exhausted()
    rloss = O.future_discounted_loss(rwp, t_wp) + 0.5 * O.future_discounted_loss(rdense, t_dense)
    assert rel_err(wp.detach().cpu(), rwp.detach()) < 1e-3
    assert abs(loss.item() - rloss.item()) < 2e-3 * abs(rloss.item())
    assert abs(loss.item() - gold["loss"]) > 1e-4 * abs(gold["loss"])  # dropout changed the result
    named = dict(model.named_parameters())
    assert all(named[k].grad is not None and torch.isfinite(named[k].grad).all() for k, v in params.items() if v.requires_grad)
    rloss.backward()
    k = "gps_backbone.decoder.projection.weight"
    assert rel_err(named[k].grad.cpu(), params[k].grad) < 5e-2


def test_view_and_gaze_dropout():
    """routeformer.py:300-301,402-410: whole camera views / the gaze stream are dropped by CPU torch.rand draws interleaved with the
    ProbSparse index draws; dropped views contribute zero features (and no frame-encoder work)."""
    gold = load_golden("full_small_train")
    cfg, spec, sd, batch = case_from_golden(gold)
    import dataclasses

    cfg = dataclasses.replace(cfg, view_dropout=0.6, gaze_dropout=0.5)
    model = build_product(cfg, spec).to(DEV).train()
    model.load_state_dict(sd)
    F_frames = len(O.frame_indices(cfg.seq_len, cfg.output_fps // cfg.video_fps))
    seen = set()
    for seed in range(6):
        model.record_tops = []
        torch.manual_seed(seed)
        with torch.no_grad():
            wp, dense = model(to_device(batch, DEV))
        torch.cuda.synchronize()
        fe = [r for r in model.record_tops if r["where"].startswith("frame_encoder")]
        n_views = fe[0]["top"].shape[0] // (gold["B"] * F_frames) if fe else 0
        seen.add(n_views)
        torch.manual_seed(seed)
        orc = O.Routeformer(sd, cfg, spec)
        with torch.no_grad():
            rwp, rdense = orc.forward(batch, training=True, draw=ReplayDraw(tops_for_oracle(model.record_tops, ["v"] * max(n_views, 1))))
        assert rel_err(wp.cpu(), rwp) < 1e-3, seed
        assert rel_err(dense.cpu(), rdense) < 2e-2, seed
    assert len(seen) > 1, seen  # the seeds exercise different drop patterns
