"""TF32 operand handling: TMA TFLOAT32 (round) vs FLOAT32 (the MMA truncates).  Run with RF_TMA_TF32_ROUND=0/1."""
import os
import torch
from routeformer_b200 import ops
g = torch.Generator().manual_seed(0)
M, N, K = 2048, 512, 1024
A, B = torch.rand(M, K, generator=g) + 0.5, torch.rand(N, K, generator=g) + 0.5   # positive operands expose a truncation bias
ref = A.double() @ B.double().t()
out = torch.empty(M, N, device="cuda")
ops.gemm(A.cuda(), B.cuda(), out)
err = out.cpu().double() - ref
print("RF_TMA_TF32_ROUND=", os.environ.get("RF_TMA_TF32_ROUND", "1"), "rel_err", (err.norm() / ref.norm()).item(),
      "mean signed rel", (err / ref).mean().item())
torch.backends.cuda.matmul.allow_tf32 = True
e2 = (A.cuda() @ B.cuda().t()).cpu().double() - ref
print("  cuBLAS tf32: rel_err", (e2.norm() / ref.norm()).item(), "mean signed rel", (e2 / ref).mean().item())
