"""One forward of the fused TFAM kernel at B = 256 (two clips per pass) and at B = 2 (latency regime) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = vmc.AMO_CLIP(device=dev).to(dev).eval()
for B in (256, 2):
    rgb, mot = torch.randn(B, 16, 512, device=dev), torch.randn(B, 15, 512, device=dev)
    for _ in range(2):
        out = m(rgb, mot)
    torch.cuda.synchronize()
    print(B, float(out.abs().mean()))
