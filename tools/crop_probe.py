"""FoV-crop problems alone (for ncu captures): the micro-benchmark's gaze window on 324x326 front frames (224^2 crops, window 0.5,
planar bf16 and patch-major fp16) and the pad-to-square 86x384 scene view."""
import sys
sys.path.insert(0, ".")
import torch
from routeformer_b200 import ops
BACKBONE_MEAN, BACKBONE_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
DEV = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
n = 1024
front = torch.rand(n, 3, 324, 326, device=DEV).half()
scene = torch.rand(n // 2, 3, 86, 384, device=DEV).half()
centers = (0.5 + 0.1 * torch.randn(n, 2, generator=g)).clamp(0, 1).to(DEV)
win = torch.full((n, 2), 0.5, device=DEV)
c2 = torch.tensor([[0.5, 0.5]]).repeat(n // 2, 1).to(DEV)
w2 = torch.tensor([[1.0, 384.0 / 86.0]]).repeat(n // 2, 1).to(DEV)
for _ in range(2):
    ops.fov_crop(front, centers, win, 224, BACKBONE_MEAN, BACKBONE_STD, patch=0, out_dtype=torch.bfloat16)
    ops.fov_crop(front, centers, win, 224, BACKBONE_MEAN, BACKBONE_STD, patch=28, out_dtype=torch.float16)
    ops.fov_crop(scene, c2, w2, 256, BACKBONE_MEAN, BACKBONE_STD, patch=32, out_dtype=torch.float16)
torch.cuda.synchronize()
print("ok")
