"""One qkv-shaped GEMM (ViT-B/16, 2048 frames in flight) for an ncu --set full capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vimoclip_b200 import ops
dev = torch.device("cuda:0")
M, d = 2048 * 197, 768
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
w = (torch.randn(3 * d, d, device=dev, generator=gen) * d**-0.5).to(torch.bfloat16)
b = torch.randn(3 * d, device=dev, generator=gen)
out = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(x, w, bias=b, out=out)
torch.cuda.synchronize()
