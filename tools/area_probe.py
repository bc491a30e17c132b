"""Area-resize problems alone (for ncu captures): the reference's GEM front frames (1080x1088 x0.3) and GoPro frames (2160x3840, rows
648..1512, x0.1)."""
import sys
sys.path.insert(0, ".")
import torch
from routeformer_b200 import ops
DEV = "cuda"
g = torch.Generator(device=DEV).manual_seed(0)
front = torch.randint(0, 256, (96, 3, 1080, 1088), device=DEV, dtype=torch.uint8, generator=g)
gopro = torch.randint(0, 256, (24, 3, 2160, 3840), device=DEV, dtype=torch.uint8, generator=g)
for _ in range(2):
    ops.area_resize_u8(front, 0.3)
    ops.area_resize_u8(gopro, 0.1, rows=(648, 1512))
torch.cuda.synchronize()
print("ok")
