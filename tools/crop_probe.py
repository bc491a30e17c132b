"""FoV-crop problems alone (for ncu captures): gaze window on front frames, and the pad-to-square scene view."""
import sys
sys.path.insert(0, ".")
import torch
from oracle.routeformer_oracle import BackboneSpec, frame_window
from routeformer_b200 import ops
DEV = "cuda"
spec = BackboneSpec()
g = torch.Generator(device="cpu").manual_seed(0)
n = 512
front = torch.rand(n, 3, 324, 326, device=DEV).half()
scene = torch.rand(n, 3, 86, 384, device=DEV).half()
centers = (0.5 + 0.1 * torch.randn(n, 2, generator=g)).clamp(0, 1).to(DEV)
win = torch.full((n, 2), 0.5, device=DEV)
cx, cy, fw, fh = frame_window(86, 384)
c2 = torch.tensor([[cx, cy]]).repeat(n, 1).to(DEV)
w2 = torch.tensor([[fw, fh]]).repeat(n, 1).to(DEV)
for od in (torch.float16, torch.float32):
    for _ in range(2):
        ops.fov_crop(front, centers, win, 256, spec.mean, spec.std, patch=32, out_dtype=od)
        ops.fov_crop(scene, c2, w2, 256, spec.mean, spec.std, patch=32, out_dtype=od)
torch.cuda.synchronize()
print("ok")
