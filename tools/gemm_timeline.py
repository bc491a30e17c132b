"""In-kernel timeline of the GEMM CTA (0,0,0) for a few shapes, plus micro timings with the TMA-store epilogue on/off."""
import sys
import torch
from routeformer_b200 import ops, _lib
lib = _lib.load()
st = torch.zeros(8, dtype=torch.int64, device="cuda")
def timeline(M, N, K, **kw):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    for _ in range(3): ops.gemm(A, B, C, **kw)
    torch.cuda.synchronize()
    lib.rf_debug_gemm_stamps(st.data_ptr())
    ops.gemm(A, B, C, **kw); torch.cuda.synchronize()
    lib.rf_debug_gemm_stamps(None)
    s = st.cpu().tolist()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gemm(A, B, C, **kw)
    e1.record(); torch.cuda.synchronize()
    names = ["setup", "first_operands", "mma_issue_done", "accum_ready", "epilogue", "exit"]
    d = [s[i + 1] - s[i] for i in range(6)]
    print(f"M={M} N={N} K={K}: kernel {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us | CTA0 ns: " + " ".join(f"{n}={v}" for n, v in zip(names, d)), flush=True)
for shp in [(2560, 128, 128), (2560, 128, 832), (99840, 128, 128), (99840, 384, 128), (99840, 256, 128), (2560, 832, 3328), (98304, 1024, 3072)]:
    timeline(*shp)
