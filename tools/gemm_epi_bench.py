"""Frame-encoder FFN GEMMs with their real epilogues (the two heaviest GEMM call sites of the step)."""
import sys
sys.path.insert(0, ".")
import torch
from routeformer_b200 import ops
DEV = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
M, D, F = 99840, 128, 256
x = torch.randn(M, D, device=DEV); w1 = torch.randn(F, D, device=DEV) / 11; b1 = torch.randn(F, device=DEV)
h = torch.empty(M, F, device=DEV); pre = torch.empty(M, F, device=DEV)
dy = torch.randn(M, D, device=DEV); w2 = torch.randn(D, F, device=DEV) / 16; dpre = torch.empty(M, F, device=DEV)
res = torch.randn(M, D, device=DEV); y = torch.empty(M, D, device=DEV); b2 = torch.randn(D, device=DEV)
gb1 = torch.zeros(F, device=DEV); gb2 = torch.zeros(D, device=DEV)
cases = {
    "ffn1 plain                 ": lambda: ops.gemm(x, w1, h),
    "ffn1 +bias                 ": lambda: ops.gemm(x, w1, h, bias=b1),
    "ffn1 +bias+gelu            ": lambda: ops.gemm(x, w1, h, bias=b1, act=ops.ACT_GELU),
    "ffn1 +bias+gelu+preact     ": lambda: ops.gemm(x, w1, h, bias=b1, act=ops.ACT_GELU, preact=pre),
    "ffn1 +bias+gelu+savegrad   ": lambda: ops.gemm(x, w1, h, bias=b1, act=ops.ACT_GELU_SAVE_GRAD, preact=pre),
    "ffn1 +bias+relu            ": lambda: ops.gemm(x, w1, h, bias=b1, act=ops.ACT_RELU),
    "ffn2 +bias+residual        ": lambda: ops.gemm(h, w2, y, bias=b2, residual=res),
    "dgrad plain (b_mn)         ": lambda: ops.gemm(dy, w2, dpre, b_mn=True),
    "dgrad +dgelu(aux)          ": lambda: ops.gemm(dy, w2, dpre, b_mn=True, dact=ops.ACT_GELU, dact_aux=pre),
    "dgrad *saved(aux)          ": lambda: ops.gemm(dy, w2, dpre, b_mn=True, dact=ops.DACT_SAVED, dact_aux=pre),
    "dgrad +drelu(aux)          ": lambda: ops.gemm(dy, w2, dpre, b_mn=True, dact=ops.ACT_RELU, dact_aux=h),
    "dgrad *saved(aux) +colsum_a": lambda: ops.gemm(dy, w2, dpre, b_mn=True, dact=ops.DACT_SAVED, dact_aux=pre, colsum_a=gb2),
    "dx dgrad +residual         ": lambda: ops.gemm(dpre, w1, y, b_mn=True, residual=res),
    "dx dgrad +residual+colsum_a": lambda: ops.gemm(dpre, w1, y, b_mn=True, residual=res, colsum_a=gb1),
    "separate colsum [M,256]    ": lambda: ops.colsum_accumulate(dpre, gb1),
}
for name, fn in cases.items():
    print(f"{name} {timeit(fn):8.1f} us", flush=True)
