"""Frame-encoder attention problem alone (for ncu captures): 1536 sequences x 8 heads, L = 65, dh = 16."""
import math
import sys
sys.path.insert(0, ".")
import torch
from routeformer_b200 import ops
DEV = "cuda"
B, H, L, dh, factor = 1536, 8, 65, 16, 5
D = H * dh
qkv = torch.randn(B * L, 3 * D, device=DEV)
U = u = min(factor * math.ceil(math.log(L)), L)
idx = torch.randint(L, (L, U), device=DEV, dtype=torch.int32)
out = torch.empty(B * L, D, device=DEV)
top = torch.empty(B, H, u, dtype=torch.int32, device=DEV)
q = (qkv, L * 3 * D, 3 * D); k = (qkv[:, D:], L * 3 * D, 3 * D); v = (qkv[:, 2 * D:], L * 3 * D, 3 * D)
dq = torch.empty_like(qkv)
for _ in range(3):
    ops.attention_fwd(q, k, v, B, H, L, L, dh, ops.ATTN_PROB, 0, idx, 0, U, u, out, top)
    ops.attention_bwd(q, k, v, B, H, L, L, dh, ops.ATTN_PROB, 0, U, u, top, out, dq, dq[:, D:], dq[:, 2 * D:])
torch.cuda.synchronize()
print("ok")
