"""Launch the HBM-bound kernels of the step at bench shapes (2048 frames in flight) for an ncu --set full capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vimoclip_b200 import ops
dev = torch.device("cuda:0")
gen = torch.Generator(device="cuda").manual_seed(0)
u8 = torch.randint(0, 256, (2048, 3, 224, 224), dtype=torch.uint8, device=dev, generator=gen)
bgr = torch.randint(0, 256, (64, 17, 224, 224, 3), dtype=torch.uint8, device=dev, generator=gen)
x = torch.randn(2048 * 197, 768, device=dev, generator=gen)
g_ = torch.ones(768, device=dev)
for _ in range(3):
    ops.prologue(u8, wrap=False, dst="patch", patch=16)
    ops.prologue(u8, wrap=True, dst="patch", patch=32)
    ops.frame_diff(bgr, dst="patch", patch=32, want_diff=False)
    ops.layernorm(x, g_, g_)
torch.cuda.synchronize()
