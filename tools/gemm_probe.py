"""The two dominant HBM-bound GEMM shapes of the frame encoder alone (for ncu source-level captures), next to plain device
copies / fills of the same footprint (what a trivial streaming kernel reaches on buffers of this size, L2 flushed)."""
import sys
sys.path.insert(0, ".")
import torch
from routeformer_b200 import ops
DEV = "cuda"
flush = torch.empty(512 * 1024 * 1024 // 4, device=DEV)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


SCALE = int(sys.argv[sys.argv.index('--scale') + 1]) if '--scale' in sys.argv else 1  # steady state: time(4M) - time(M)
M, D = 99840 * SCALE, 128
x = torch.randn(M, D, device=DEV)
wqkv = torch.randn(3 * D, D, device=DEV) / 11; bqkv = torch.randn(3 * D, device=DEV); qkv = torch.empty(M, 3 * D, device=DEV)
w1 = torch.randn(2 * D, D, device=DEV) / 11; h = torch.empty(M, 2 * D, device=DEV)
wo = torch.randn(D, D, device=DEV) / 11; y = torch.empty(M, D, device=DEV)
src3 = torch.randn(M, 3 * D, device=DEV)
rows = [
    ("qkv   N=384 K=128 +bias", lambda: ops.gemm(x, wqkv, qkv, bias=bqkv), 4.0 * (M * D + M * 3 * D)),
    ("ffn1  N=256 K=128 plain", lambda: ops.gemm(x, w1, h), 4.0 * (M * D + M * 2 * D)),
    ("out   N=128 K=128 plain", lambda: ops.gemm(x, wo, y), 4.0 * (M * D + M * D)),
    ("dqkv  N=128 K=384 b_mn ", lambda: ops.gemm(src3, wqkv, y, b_mn=True), 4.0 * (M * 3 * D + M * D)),
    ("fill  [M,384]          ", lambda: qkv.zero_(), 4.0 * M * 3 * D),
    ("copy  [M,128]->[M,128] ", lambda: y.copy_(x), 4.0 * 2 * M * D),
    ("copy  [M,384]->[M,384] ", lambda: qkv.copy_(src3), 4.0 * 2 * M * 3 * D),
]
if "--chain" in sys.argv:  # back-to-back launches (no flush in between, PDL edges as inside the step): per-launch time in a chain
    for name, fn, nbytes in rows:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / 20
        print(f"{name} chained {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s", flush=True)
    sys.exit(0)
if "--stages" in sys.argv:  # which pipeline stage bounds the persistent kernel?  (probe bits: see gemm_tf32.cu g_probe)
    from routeformer_b200 import _lib
    lib = _lib.load()
    modes = [(0, "normal"), (1, "no bulk store"), (2, "epilogue drains TMEM only"), (4, "B loaded once per CTA"), (8, "no MMAs"),
             (2 | 8, "loads only (no MMAs, no epilogue work)"), (2 | 4 | 8, "A loads only"), (1 | 4, "no store, B once")]
    print("shape".ljust(26) + "".join(f"{n[:22]:>24s}" for _, n in modes))
    for name, fn, nbytes in rows[:4]:
        line = name.ljust(26)
        for m, _ in modes:
            lib.rf_debug_gemm_probe(m)
            line += f"{timeit(fn):21.1f} us"
        lib.rf_debug_gemm_probe(0)
        print(line, flush=True)
    sys.exit(0)
if "--ncu" in sys.argv:  # two launches of every GEMM shape, nothing else
    for name, fn, nbytes in rows[:4]:
        fn(); flush.zero_(); fn()
    torch.cuda.synchronize()
    sys.exit(0)
for name, fn, nbytes in rows:
    us = timeit(fn)
    print(f"{name} {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s", flush=True)
