#!/usr/bin/env python
"""Routeformer fwd+bwd training-step benchmark (BASELINE.json metric: clips/s at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                   (the reference algorithm's CPU path on this box's host cores)

One "step" = forward + loss + backward + gradient all-reduce + clip + AdamW of the full-modality Routeformer (GPS + scene
video + gaze FoV, paper configuration of experiments/full_comparison.py:159-296 with the build-defined random-init patch
backbone, dropouts 0) on a synthetic GEM-shaped batch of 64 clips per GPU (weak scaling: 512 clips on 8 GPUs).
`value` times the step with the batch resident in HBM; `e2e` times the same step through the public API from pinned host
buffers (consumed frames staged H2D every step, loss read back every step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAPER = dict(with_video=True, with_gaze=True, dense_prediction=True)  # + OracleConfig defaults = paper configuration
BACKBONE = dict(image_size=256, patch=32, channels=1024, window=0.5)


def build_case(B: int, seed: int, fov: str):
    from oracle import routeformer_oracle as O  # synthetic-batch / config description only (shared with the tests)

    cfg = O.OracleConfig(**PAPER)
    spec = O.BackboneSpec(**BACKBONE)
    batch = O.synthetic_batch(B, cfg, "gem", seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    targets = (batch["gps"][:, -1:, :] + torch.cumsum(1.83 + 0.91 * torch.randn(B, cfg.pred_len, 2, generator=g), 1),
               torch.randn(B, cfg.pred_len, cfg.image_embedding_size, generator=g))
    return cfg, spec, batch, targets


def build_model(cfg, spec, fov):
    import routeformer_b200 as R

    g = R.GPSBackboneConfig(seq_len=cfg.seq_len, label_len=cfg.seq_len, pred_len=cfg.pred_len, factor=cfg.factor, distil=True,
                            dropout=0.0, activation="relu", d_model=cfg.d_model, n_heads=cfg.n_heads, e_layers=cfg.e_layers,
                            d_layers=cfg.d_layers, d_ff=cfg.d_ff)
    vb = R.PatchBackboneConfig(image_size=spec.image_size, patch=spec.patch, channels=spec.channels, fov=fov, window=spec.window)
    rc = R.RouteformerConfig(gps_backbone_config=g, video_backbone_config=vb, decoder_mode="smart", with_video=True, with_gaze=True,
                             dense_prediction=True, dense_loss_ratio=0.5, image_embedding_size=64, encoder_hidden_size=64,
                             encoder_heads=8, encoder_layers=8, encoder_d_ff=256, cross_modal_decoder_heads=8,
                             cross_modal_decoder_layers=2, view_dropout=0.0, gaze_dropout=0.0, feature_dropout=0.0)
    return R.Routeformer(rc, gps_backbone=R.Informer, video_backbone=R.PatchEmbedBackbone)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(2)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import routeformer_b200 as R
    from routeformer_b200 import ops
    from routeformer_b200.parallel import DataParallelTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))  # a stuck collective aborts quickly
    B = args.batch_per_gpu
    cfg, spec, host_batch, host_targets = build_case(B, seed=100 + rank, fov=args.fov)
    torch.manual_seed(0)
    model = build_model(cfg, spec, args.fov).to(dev).train()
    if args.no_branch_overlap:
        R.Routeformer.overlap_branches = False
    lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")

    def loss_fn(out, tgt):
        wp, dense = out
        return lossf(wp, tgt[0]) + 0.5 * lossf(dense, tgt[1])

    trainer = DataParallelTrainer(model, loss_fn, lr=1e-5, weight_decay=1e-4, max_grad_norm=2.5, use_cuda_graph=not args.no_graph,
                                  overlap_wgrad=not args.no_wgrad_overlap)
    trainer.broadcast_parameters()
    # inputs: pinned host batch in the reference layout; the device batch holds the frames the model consumes (8 of 40 per view)
    pinned = {k: v.contiguous().pin_memory() for k, v in host_batch.items()}
    pinned_t = tuple(t.contiguous().pin_memory() for t in host_targets)
    batch = model.stage_batch(pinned, dev)
    targets = tuple(t.to(dev) for t in host_targets)
    in_bytes = sum(v.numel() * v.element_size() for v in batch.values())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(args.warmup):
        trainer.step(batch, targets)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = trainer.step(batch, targets)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (ops.launch_count - launches0) // max(args.steps, 1) + trainer.graph_launches
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "value": round(value, 2), "ms_per_step": round(ms / args.steps, 3), "gpu_launches": int(launches)}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public API: pinned host batch -> stage -> step -> loss.item() ----
    e2e_steps = max(2, min(args.steps, 5))
    h2d = in_bytes + sum(t.numel() * t.element_size() for t in pinned_t)

    from routeformer_b200.parallel import BatchPrefetcher
    prefetch = BatchPrefetcher(model, dev, depth=2)

    def e2e_step():
        # public-API pipeline: the H2D staging of the NEXT step's inputs (consumed frames + gps + gaze + targets, from pinned
        # host memory, every step) runs on a side stream while this step computes; the loss is read back every step.
        # Order matters: the step is enqueued first so that its own tiny H2D copy (the index tables, first node of the graph)
        # does not queue behind the 529 MB staging transfer on the single H2D copy engine.
        b, t = prefetch.get()
        loss_t = trainer.step(b, t)
        prefetch.release(b)
        prefetch.submit(pinned, pinned_t)
        trainer.prefetch_draws()  # host RNG work of the next step, hidden behind this step's device time
        return loss_t.item()

    prefetch.submit(pinned, pinned_t)
    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        loss_host = e2e_step()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(ms_e2e.item()) / 1e3)

    # ---- roofline of the dominant kernel family (tcgen05 GEMM): one instrumented step, CUDA events per launch ----
    roofline = cpu_baseline = None
    kernels = {}
    # every rank runs the instrumented step (it contains the gradient all-reduce); only rank 0 reports
    if True:
        peaks, peak_src = measured_peaks()
        rec = []
        orig_gemm, orig_crop, orig_af, orig_ab = ops.gemm, ops.fov_crop, ops.attention_fwd, ops.attention_bwd

        gemm_calls = []  # (args, kwargs, work) of every GEMM launch of the step, replayed below for clean per-launch timings

        def timed(fn, tag, work):
            def wrapper(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = fn(*a, **k)
                e.record()
                w = work(*a, **k)
                rec.append((tag, s, e, w))
                if tag == "gemm":
                    gemm_calls.append((a, k, w))
                return r
            return wrapper

        shapes = {}

        def gemm_work(A, Bm, out, **k):
            Kd = A.shape[0] if k.get("a_mn") else A.shape[1]
            Md = A.shape[1] if k.get("a_mn") else A.shape[0]
            Nd = Bm.shape[1] if k.get("b_mn") else Bm.shape[0]
            key = (Md, Nd, Kd, bool(k.get("a_mn")), bool(k.get("b_mn")), bool(k.get("accumulate")))
            shapes.setdefault(key, []).append(len(rec))
            # algorithmic bytes: both operands once, the output once, plus every [M,N] epilogue operand / second output
            extra = sum(1 for name in ("residual", "preact", "dact_aux") if k.get(name) is not None) + (1 if k.get("accumulate") else 0)
            return {"flops": 2.0 * Md * Nd * Kd, "bytes": 4.0 * (Md * Kd + Nd * Kd + (1 + extra) * Md * Nd)}

        def crop_work(frames, centers, windows, S, mean, std, **k):
            n = centers.shape[0]
            H, W = frames.shape[-2:]
            win = windows[0].tolist()
            src = 3 * min(H, win[1] * H + 2) * min(W, win[0] * W + 2) * frames.element_size()
            out_t = k.get("out")
            return n * (src + 3 * S * S * (out_t.element_size() if out_t is not None else 4))

        ops.gemm = timed(orig_gemm, "gemm", gemm_work)
        ops.fov_crop = timed(orig_crop, "fov_crop", crop_work)
        ops.attention_fwd = timed(orig_af, "attention", lambda *a, **k: 0.0)
        ops.attention_bwd = timed(orig_ab, "attention", lambda *a, **k: 0.0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        graph_mode, trainer.use_cuda_graph = trainer.use_cuda_graph, False  # per-launch events need the eager path
        trainer.step(batch, targets)
        trainer.use_cuda_graph = graph_mode
        ev1.record()
        torch.cuda.synchronize()
        ops.gemm, ops.fov_crop, ops.attention_fwd, ops.attention_bwd = orig_gemm, orig_crop, orig_af, orig_ab
        step_ms = ev0.elapsed_time(ev1)
        hbm_peak = peaks["hbm_gbs"]
        tf32_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2.0  # kind::tf32 runs at half the bf16 rate
        for tag in ("fov_crop", "attention"):
            rows = [(s.elapsed_time(e), w) for t_, s, e, w in rec if t_ == tag]
            if rows:
                # event pairs around eager launches: exact for the long kernels (crop), inflated by host issue time for short ones
                kernels[tag] = {"launches": len(rows), "ms": round(sum(r[0] for r in rows), 3), "work": sum(r[1] for r in rows)}
                if tag == "attention":
                    kernels[tag]["ms_note"] = "eager event pairs, includes host launch gaps; see profiles/*ncu_launch_summary* for device times"
        # GEMM launches, split by the roofline that bounds each one (time at peak: bytes / HBM vs flops / TF32).  The eager step
        # is host-bound (an event pair then also measures the ~10-20 us the host needs to issue the launch), so the launches that
        # matter -- at least 8 us at the roofline -- are REPLAYED with their own arguments behind a queue of work that keeps the
        # host ahead of the device; each one is timed by its own event pair on the launching stream.
        g_rows = [(s.elapsed_time(e), w) for t_, s, e, w in rec if t_ == "gemm"]
        at_peak = lambda w: max(w["bytes"] / (hbm_peak * 1e9), w["flops"] / (tf32_peak * 1e12))
        replay = [(a_, k_, w) for a_, k_, w in gemm_calls if at_peak(w) >= 8e-6]
        blocker = torch.empty(256 << 20, device=dev, dtype=torch.float32)
        timings = []
        for _ in range(2):  # first pass warms the tensor-map / attribute caches
            timings.clear()
            for _ in range(4):
                blocker.zero_()
            for a_, k_, w in replay:
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                orig_gemm(*a_, **k_)
                e_.record()
                timings.append((s_, e_, w))
            torch.cuda.synchronize()
        del blocker
        fam = {"hbm": [0.0, 0.0, 0.0, 0], "tensor": [0.0, 0.0, 0.0, 0]}
        for s_, e_, w in timings:
            ms_l = s_.elapsed_time(e_)
            f = fam["hbm" if w["bytes"] / (hbm_peak * 1e9) >= w["flops"] / (tf32_peak * 1e12) else "tensor"]
            f[0] += ms_l; f[1] += w["bytes"]; f[2] += w["flops"]; f[3] += 1
        gemm_calls.clear()
        kernels["gemm"] = {"launches": len(g_rows), "replayed_for_roofline": len(timings)}
        for name, (ms_f, by, fl, n) in fam.items():
            if n:
                kernels[f"gemm_{name}_bound"] = {"launches": n, "ms": round(ms_f, 3), "achieved_gbs": round(by / ms_f / 1e6, 1),
                                                 "achieved_tflops": round(fl / ms_f / 1e9, 1), "avg_launch_us": round(1e3 * ms_f / n, 2)}
        dom = "hbm" if fam["hbm"][0] >= fam["tensor"][0] else "tensor"
        ms_f, by, fl, n = fam[dom]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")  # per-launch dram bytes of the same launches (ncu --set full)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        if dom == "hbm":
            achieved, peak, unit = by / ms_f / 1e6, hbm_peak, "GB/s"
            src = f"{peak_src} HBM copy bandwidth"
        else:
            achieved, peak, unit = fl / ms_f / 1e9, tf32_peak, "TFLOP/s"
            src = f"{peak_src} bf16 dense sustained / 2 (kind::tf32)"
        roofline = {"kernel": "gemm_tf32_persistent_kernel / gemm_tf32_kernel (tcgen05.mma kind::tf32): the HBM-bound launches of one step "
                              "(frame-encoder QKV / FFN / projection GEMMs, K = 128..384)" if dom == "hbm" else
                              "gemm_tf32_kernel (tcgen05.mma kind::tf32): the tensor-bound launches of one step",
                    "bound": dom, "achieved": round(achieved, 1), "peak": peak, "unit": unit, "frac": round(achieved / peak, 4),
                    "traffic": traffic, "peak_source": src, "launches": n, "avg_launch_us": round(1e3 * ms_f / n, 2),
                    "algorithmic_bytes_per_launch": round(by / n), "algorithmic_flops_per_launch": round(fl / n),
                    "share_of_step": round(ms_f / (ms / args.steps), 3),  # replayed launch time / timed step (compare with the ncu list)
                    "timing": "the step's own launches (same tensors and arguments), replayed back to back after the step with one CUDA "
                              "event pair per launch on the launching stream; launches below 8 us at the roofline are left out"}
        if "fov_crop" in kernels:
            c = kernels["fov_crop"]
            gbs = c["work"] / (c["ms"] / 1e3) / 1e9
            kernels["fov_crop"].update({"achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4)})
        for k in kernels.values():
            k.pop("work", None)
        if rank == 0 and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
            with open(os.path.join(ROOT, "gpurun_out", "gemm_shapes.txt"), "w") as f:
                table = []
                for key, idxs in shapes.items():
                    ms_list = [rec[i][1].elapsed_time(rec[i][2]) for i in idxs]
                    table.append((sum(ms_list), len(idxs), key))
                for tot, cnt, (Md, Nd, Kd, amn, bmn, accf) in sorted(table, reverse=True):
                    fl = 2.0 * Md * Nd * Kd
                    by = 4.0 * (Md * Kd + Nd * Kd + Md * Nd)  # without epilogue operands
                    f.write(f"{tot:8.3f} ms {cnt:3d} x {1e3 * tot / cnt:8.1f} us  M={Md:6d} N={Nd:5d} K={Kd:6d} a_mn={int(amn)} b_mn={int(bmn)} acc={int(accf)}"
                            f"  {fl / (tot / cnt) / 1e9:7.1f} TFLOP/s {by / (tot / cnt) / 1e6:7.0f} GB/s\n")
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cpu_baseline = cpu_reference(cfg, spec, sample_clips=8, steps=2, warmup=1)
    barrier()

    if rank == 0:
        line = {
            "metric": "routeformer_fwd_bwd_clips_per_sec", "value": round(value, 2), "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (fp32 storage, fp32 accumulate; fp16 operands in the patch embedding like the reference's autocast backbone)", "data": "synthetic",
            "config": {"workload": "Routeformer GPS+scene video+gaze FoV training step (fwd+loss+bwd+allreduce+clip+AdamW), "
                                   "paper config, random-init patch backbone 256^2/p32/C1024, GEM-shaped clips",
                       "global_batch": world * B, "batch_per_gpu": B, "parallelism": f"dp{world}", "fov": args.fov,
                       "l2": f"inputs larger than L2 ({in_bytes / 1e6:.0f} MB of consumed frames + gps + gaze per step per GPU)",
                       "cuda_graph": bool(trainer.use_cuda_graph),
                       "dropout": "feature/view/gaze dropout 0 (parity configuration)", "backbone": "frozen (reference: epoch <= 10)"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "clips/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "loss": loss_host},
            "gpu_launches": int(launches),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline,
            "loss": float(loss.item()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def cpu_reference(cfg, spec, sample_clips: int, steps: int, warmup: int):
    """The reference algorithm's CPU path (oracle port: functional PyTorch restatement pinned against the reference) timed on
    the host cores: fwd + loss + bwd + AdamW on a bounded sample of the same workload."""
    from oracle import routeformer_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.fill_state_dict(O.state_dict_template(cfg, spec), 0)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe") and
                                          not k.startswith("video_backbone")) for k, v in sd.items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1e-5, weight_decay=1e-4)
    batch = O.synthetic_batch(sample_clips, cfg, "gem", seed=5)
    g = torch.Generator().manual_seed(6)
    t_wp = batch["gps"][:, -1:, :] + torch.cumsum(1.83 + 0.91 * torch.randn(sample_clips, cfg.pred_len, 2, generator=g), 1)
    t_dense = torch.randn(sample_clips, cfg.pred_len, cfg.image_embedding_size, generator=g)
    model = O.Routeformer(params, cfg, spec)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        wp, dense = model.forward(batch, training=True)
        loss = O.future_discounted_loss(wp, t_wp) + 0.5 * O.future_discounted_loss(dense, t_dense)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.requires_grad], 2.5)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return {"value": round(sample_clips / sec, 3), "unit": "clips/s", "cores": cores, "kind": "port",
            "sample": f"{sample_clips} GEM-shaped clips per step, {steps} timed steps after {warmup} warm-up, fp32, torch {torch.__version__} CPU",
            "sec_per_step": round(sec, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import routeformer_oracle as O

    cfg, spec = O.OracleConfig(**PAPER), O.BackboneSpec(**BACKBONE)
    sample = 8
    # K and W are honoured up to a bound that keeps the CPU run within a few minutes (one step of the 8-clip sample takes ~0.7 s)
    steps, warmup = max(1, min(args.steps, 200)), max(1, min(args.warmup, 10))
    res = cpu_reference(cfg, spec, sample_clips=sample, steps=steps, warmup=warmup)
    line = {
        "impl": "reference", "metric": "routeformer_fwd_bwd_clips_per_sec", "value": res["value"], "unit": "clips/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(1e3 * res["sec_per_step"], 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Routeformer GPS+scene video+gaze FoV training step, paper config, CPU (reference algorithm, oracle port)",
                   "sample_clips_per_step": sample},
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)  # ~0.8 s timed region: several nvidia-smi clock samples fall inside it
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=64)
    ap.add_argument("--fov", default="gaze", choices=["gaze", "frame"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-branch-overlap", action="store_true", help="run the gaze encoder on the main stream instead of a side stream")
    ap.add_argument("--no-wgrad-overlap", action="store_true", help="keep the weight-gradient GEMMs on the main stream")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying one captured CUDA graph")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: 1 warm-up + --steps steps, no e2e / roofline / CPU legs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not args.profile:
        args.warmup = 3
    if args.profile:
        args.warmup = 1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
