"""Attention forward micro-benchmark: tcgen05 kernel vs the fp32 FMA kernels on the frame-encoder / gaze-encoder problems.
usage (GPU box): python tools/attn_bench.py  -> prints one line per (shape, kernel) with the mean launch time in microseconds."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from routeformer_b200 import ops  # noqa: E402


def run(B, H, L, dh, factor, reps=20):
    import math
    D = H * dh
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(L)
    qkv = torch.randn(B * L, 3 * D, device=dev, generator=g)
    U = u = min(L, factor * int(math.ceil(math.log(L))))
    idx = torch.randint(L, (1, L, U), device=dev, generator=g, dtype=torch.int32)
    q = (qkv, L * 3 * D, 3 * D)
    k = (qkv[:, D:], L * 3 * D, 3 * D)
    v = (qkv[:, 2 * D:], L * 3 * D, 3 * D)
    outs = {}
    for tc in ("0", "1"):
        os.environ["RF_ATTN_TC"] = tc
        out = torch.empty(B * L, D, device=dev)
        top = torch.zeros(B, H, u, dtype=torch.int32, device=dev)
        for _ in range(3):
            ops.attention_fwd(q, k, v, B, H, L, L, dh, ops.ATTN_PROB, ops.LAYOUT_BLHD, idx, 0, U, u, out, top)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.attention_fwd(q, k, v, B, H, L, L, dh, ops.ATTN_PROB, ops.LAYOUT_BLHD, idx, 0, U, u, out, top)
        e1.record()
        torch.cuda.synchronize()
        outs[tc] = (out.clone(), top.clone())
        print(f"B={B} H={H} L={L} dh={dh} kernel={'tcgen05' if tc == '1' else 'fp32 FMA'}: {1e3 * e0.elapsed_time(e1) / reps:8.1f} us / launch")
    same = torch.equal(outs["0"][1].sort(-1).values, outs["1"][1].sort(-1).values)
    err = ((outs["0"][0] - outs["1"][0]).norm() / outs["0"][0].norm()).item()
    print(f"    selections identical: {same}; context rel diff tcgen05 vs fp32: {err:.2e}")


def run_generic(B, H, L, dh, factor, mode, layout, reps=20):
    """Generic kernels (video encoder L = 160; Informer dh = 104): 128 vs 256 threads per CTA, forward and backward."""
    import math
    D = H * dh
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(L + dh)
    qkv = torch.randn(B * L, 3 * D, device=dev, generator=g)
    U = u = min(L, factor * int(math.ceil(math.log(L))))
    idx = torch.randint(L, (1, L, U), device=dev, generator=g, dtype=torch.int32)
    q, k, v = (qkv, L * 3 * D, 3 * D), (qkv[:, D:], L * 3 * D, 3 * D), (qkv[:, 2 * D:], L * 3 * D, 3 * D)
    dout = torch.randn(B * L, D, device=dev, generator=g)
    res = {}
    for wide in ("0", "1"):
        os.environ["RF_ATTN_WIDE"] = wide
        out = torch.empty(B * L, D, device=dev)
        top = torch.zeros(B, H, u, dtype=torch.int32, device=dev)
        dqkv = torch.zeros_like(qkv)
        fwd = lambda: ops.attention_fwd(q, k, v, B, H, L, L, dh, mode, layout, idx, 0, U, u, out, top)
        bwd = lambda: ops.attention_bwd(q, k, v, B, H, L, L, dh, mode, layout, U, u, top, dout, dqkv, dqkv[:, D:], dqkv[:, 2 * D:])
        t = []
        for fn in (fwd, bwd):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t.append(1e3 * e0.elapsed_time(e1) / reps)
        res[wide] = (out.clone(), dqkv.clone())
        print(f"B={B} H={H} L={L} dh={dh} mode={mode} threads={'256' if wide == '1' else '128'}: fwd {t[0]:7.1f} us  bwd {t[1]:7.1f} us")
    print(f"    identical results: out {torch.equal(res['0'][0], res['1'][0])}, grads rel diff "
          f"{((res['0'][1] - res['1'][1]).norm() / res['0'][1].norm()).item():.1e}")
    os.environ.pop("RF_ATTN_WIDE", None)


def timeline(B, H, L, dh, factor):
    import math
    from routeformer_b200 import _lib
    lib = _lib.load()
    D = H * dh
    dev = "cuda"
    qkv = torch.randn(B * L, 3 * D, device=dev)
    U = u = min(L, factor * int(math.ceil(math.log(L))))
    idx = torch.randint(L, (1, L, U), device=dev, dtype=torch.int32)
    q, k, v = (qkv, L * 3 * D, 3 * D), (qkv[:, D:], L * 3 * D, 3 * D), (qkv[:, 2 * D:], L * 3 * D, 3 * D)
    out = torch.empty(B * L, D, device=dev)
    top = torch.zeros(B, H, u, dtype=torch.int32, device=dev)
    os.environ["RF_ATTN_TC"] = "1"
    st = torch.zeros(16, dtype=torch.int64, device=dev)
    ops.attention_fwd(q, k, v, B, H, L, L, dh, ops.ATTN_PROB, ops.LAYOUT_BLHD, idx, 0, U, u, out, top)
    torch.cuda.synchronize()
    lib.rf_debug_attn_stamps(st.data_ptr())
    ops.attention_fwd(q, k, v, B, H, L, L, dh, ops.ATTN_PROB, ops.LAYOUT_BLHD, idx, 0, U, u, out, top)
    torch.cuda.synchronize()
    lib.rf_debug_attn_stamps(None)
    s = st.cpu().tolist()
    names = ["start", "setup", "counts", "tiles landed", "operands", "scores", "h0 measure", "h0 softmax", "h0 P written", "h0 PV issued",
             "h1 measure", "h1 softmax", "h1 P written", "h1 PV issued", "contexts", "exit"]
    print(f"timeline of the middle CTA (B={B} L={L}), cycles since start:")
    for n, t in zip(names, s):
        print(f"    {n:14s} {t - s[0]:8d}")


if __name__ == "__main__":
    if "--generic" in sys.argv:
        run_generic(64, 8, 160, 16, 5, ops.ATTN_PROB, ops.LAYOUT_BLHD)          # video encoder
        run_generic(64, 8, 40, 104, 4, ops.ATTN_PROB, ops.LAYOUT_BHLD)          # Informer encoder layer 0
        run_generic(64, 8, 70, 104, 4, ops.ATTN_PROB_MASKED, ops.LAYOUT_BHLD)   # Informer decoder self-attention
        run_generic(64, 8, 40, 8, 5, ops.ATTN_PROB_MASKED, ops.LAYOUT_BLHD)     # gaze-video decoder self-attention
        run_generic(64, 8, 21, 104, 4, ops.ATTN_PROB, ops.LAYOUT_BHLD)          # Informer encoder layer 1
        sys.exit(0)
    timeline(1536, 8, 65, 16, 5)
    run(1536, 8, 65, 16, 5)
    run(64, 8, 40, 16, 5)
    run(512, 8, 65, 16, 5)
