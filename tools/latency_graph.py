"""Latency of the launch-bound regime, eager launches vs CUDA-graph replay (vimoclip_b200.graphed):
one clip at a time through the student (inference.py:129), a config-1 TFAM batch, one clip through the whole pipeline."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402

dev = torch.device("cuda:0")


def lat(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)
for name in ("ViT-B/32", "ViT-B/16"):
    student = vmc.FlowStudentModel(name, device=dev, num_classes=140).eval()
    clip = torch.randint(0, 256, (1, 15, 3, 224, 224), dtype=torch.uint8, generator=gen).to(dev)
    g = vmc.graphed(student, clip)
    print(f"student {name}, 1 clip x 15 frames: eager launches {lat(lambda: student(clip)):.3f} ms, graph replay {lat(lambda: g(clip)):.3f} ms", flush=True)
tfam = vmc.AMO_CLIP(num_classes=140, device=dev).to(dev).eval()
rgb, mot = torch.randn(2, 16, 512, generator=gen).to(dev), torch.randn(2, 15, 512, generator=gen).to(dev)
mr = (torch.arange(16)[None, :] < torch.tensor([16, 12])[:, None]).to(dev)
mm = (torch.arange(15)[None, :] < torch.tensor([15, 11])[:, None]).to(dev)
g = vmc.graphed(tfam, rgb, mot, mr, mm)
print(f"TFAM config 1 (B=2, T=16/15, ragged masks): eager launches {lat(lambda: tfam(rgb, mot, mr, mm)):.3f} ms, graph replay {lat(lambda: g(rgb, mot, mr, mm)):.3f} ms", flush=True)
pipe = vmc.ViMoCLIPPipeline("openai/clip-vit-base-patch16", "ViT-B/32", num_classes=140, device=dev, clips_per_step=1)
r8 = torch.randint(0, 256, (1, 16, 3, 224, 224), dtype=torch.uint8, generator=gen).to(dev)
m8 = torch.randint(0, 256, (1, 15, 3, 224, 224), dtype=torch.uint8, generator=gen).to(dev)
g = vmc.graphed(pipe, r8, m8)
print(f"full pipeline (CLIP B/16 + student B/32 + TFAM), 1 clip: eager launches {lat(lambda: pipe(r8, m8)):.3f} ms, graph replay {lat(lambda: g(r8, m8)):.3f} ms", flush=True)
