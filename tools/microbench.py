"""Kernel micro-benchmarks on one B200 (CUDA events, warm-up, L2 flushed between iterations).  Prints one line per case."""
import json
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from routeformer_b200 import ops  # noqa: E402

DEV = "cuda"
PEAKS = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_case(name, M, N, K, a_mn=False, b_mn=False, accumulate=False):
    A = torch.randn((K, M) if a_mn else (M, K), device=DEV)
    B = torch.randn((K, N) if b_mn else (N, K), device=DEV)
    C = torch.zeros(M, N, device=DEV)
    ms = timeit(lambda: ops.gemm(A, B, C, a_mn=a_mn, b_mn=b_mn, accumulate=accumulate))
    tf = 2.0 * M * N * K / ms / 1e9
    gb = 4.0 * (M * K + N * K + M * N) / ms / 1e6
    ms_t = timeit(lambda: torch.matmul(A.t() if a_mn else A, B if b_mn else B.t(), out=C)) if not accumulate else float("nan")
    print(f"gemm {name:28s} M={M:7d} N={N:5d} K={K:5d} {ms*1e3:9.1f} us  {tf:7.1f} TFLOP/s  {gb:7.0f} GB/s   torch(tf32={torch.backends.cuda.matmul.allow_tf32}) {ms_t*1e3:9.1f} us", flush=True)


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    print("peaks", PEAKS.get("hbm_gbs"), PEAKS.get("bf16_tflops"))
    Nf = 64 * 8 * 3
    gemm_case("patch_embed", Nf * 64, 1024, 3072)
    gemm_case("frame_tokenconv", Nf * 65, 384, 1024)
    gemm_case("frame_qkv", Nf * 65, 384, 128)
    gemm_case("frame_out", Nf * 65, 128, 128)
    gemm_case("frame_ffn1", Nf * 65, 256, 128)
    gemm_case("frame_ffn2", Nf * 65, 128, 256)
    gemm_case("frame_qkv_dgrad", Nf * 65, 128, 384, b_mn=True)
    gemm_case("frame_qkv_wgrad", 384, 128, Nf * 65, a_mn=True, b_mn=True, accumulate=True)
    gemm_case("frame_tokenconv_wgrad", 384, 1024, Nf * 65, a_mn=True, b_mn=True, accumulate=True)
    gemm_case("informer_qkv", 64 * 40, 2496, 832)
    gemm_case("informer_ffn1", 64 * 40, 3328, 832)
    gemm_case("informer_ffn2", 64 * 40, 832, 3328)
    gemm_case("informer_ffn1_wgrad", 3328, 832, 64 * 40, a_mn=True, b_mn=True, accumulate=True)
    gemm_case("informer_distil", 64 * 40, 2496, 832)
    gemm_case("square_8192", 8192, 8192, 8192)
    # FoV crop, config 5: 16-frame clips, 324x326 fp16 front frames -> 224^2, window 0.5, gaze jitter
    from oracle.routeformer_oracle import BackboneSpec
    spec = BackboneSpec()
    for (n, H, W, S, p, od) in [(64 * 16, 324, 326, 224, 28, torch.bfloat16), (64 * 16, 324, 326, 224, 28, torch.float32),
                                (64 * 16, 1080, 1088, 224, 28, torch.bfloat16), (64 * 8, 86, 384, 256, 32, torch.float32)]:
        frames = torch.rand(n, 3, H, W, device=DEV).half()
        g = torch.Generator(device="cpu").manual_seed(0)
        centers = (0.5 + 0.1 * torch.randn(n, 2, generator=g)).clamp(0, 1).to(DEV)
        windows = torch.full((n, 2), 0.5, device=DEV)
        out = ops.fov_crop(frames, centers, windows, S, spec.mean, spec.std, patch=p, out_dtype=od)
        ms = timeit(lambda: ops.fov_crop(frames, centers, windows, S, spec.mean, spec.std, patch=p, out_dtype=od, out=out))
        src_bytes = n * 3 * min(H, 0.5 * H + 2) * min(W, 0.5 * W + 2) * 2
        dst_bytes = out.numel() * out.element_size()
        gbs = (src_bytes + dst_bytes) / ms / 1e6
        print(f"fov_crop n={n} {H}x{W}->{S} {od}: {ms*1e3:9.1f} us  algorithmic {gbs:7.0f} GB/s = {gbs/PEAKS['hbm_gbs']:.3f} of HBM peak", flush=True)
    # attention
    for (B, H, L, dh, factor, mode, lay) in [(Nf, 8, 65, 16, 5, ops.ATTN_PROB, 0), (64, 8, 160, 16, 5, ops.ATTN_PROB, 0),
                                             (64, 8, 70, 104, 4, ops.ATTN_PROB_MASKED, 1)]:
        import math
        D = H * dh
        qkv = torch.randn(B * L, 3 * D, device=DEV)
        U = u = min(factor * math.ceil(math.log(L)), L)
        idx = torch.randint(L, (L, U), device=DEV, dtype=torch.int32)
        out = torch.empty(B * L, D, device=DEV)
        top = torch.empty(B, H, u, dtype=torch.int32, device=DEV)
        q = (qkv, L * 3 * D, 3 * D); k = (qkv[:, D:], L * 3 * D, 3 * D); v = (qkv[:, 2 * D:], L * 3 * D, 3 * D)
        ms = timeit(lambda: ops.attention_fwd(q, k, v, B, H, L, L, dh, mode, lay, idx, 0, U, u, out, top))
        dq = torch.empty_like(qkv)
        ms_b = timeit(lambda: ops.attention_bwd(q, k, v, B, H, L, L, dh, mode, lay, U, u, top, out, dq, dq[:, D:], dq[:, 2 * D:]))
        flops = B * H * (2 * L * U * dh + 4 * u * L * dh)
        print(f"attention B={B} H={H} L={L} dh={dh}: fwd {ms*1e3:8.1f} us ({flops/ms/1e9:6.2f} TFLOP/s algorithmic, {qkv.numel()*4*4/3/ms/1e6:6.0f} GB/s) bwd {ms_b*1e3:8.1f} us", flush=True)
    # layernorm
    for (M, D) in [(Nf * 65, 128), (64 * 40, 832)]:
        x = torch.randn(M, D, device=DEV); gm = torch.ones(D, device=DEV); bt = torch.zeros(D, device=DEV)
        y = torch.empty_like(x); mean = torch.empty(M, device=DEV); rstd = torch.empty(M, device=DEV)
        ms = timeit(lambda: ops.layernorm_fwd(x, gm, bt, y, mean, rstd))
        dg = torch.zeros(D, device=DEV); db = torch.zeros(D, device=DEV)
        ms_b = timeit(lambda: ops.layernorm_bwd(y, x, gm, mean, rstd, y, dg, db))
        print(f"layernorm M={M} D={D}: fwd {ms*1e3:8.1f} us ({2*x.numel()*4/ms/1e6:6.0f} GB/s) bwd {ms_b*1e3:8.1f} us ({3*x.numel()*4/ms_b/1e6:6.0f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
