"""A/B of the ViT tower LayerNorm modes inside one process (box-to-box variance is larger than the effect)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc
from vimoclip_b200 import ops, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
pipe = vmc.ViMoCLIPPipeline("openai/clip-vit-base-patch16", "ViT-B/32", num_classes=140, device=dev, clips_per_step=128)
gen = torch.Generator(device=dev).manual_seed(1)
rgb = torch.randint(0, 256, (256, 16, 3, 224, 224), dtype=torch.uint8, device=dev, generator=gen)
mot = torch.randint(0, 256, (256, 15, 3, 224, 224), dtype=torch.uint8, device=dev, generator=gen)
modes = [int(m) for m in os.environ.get("AB_MODES", "4,5,3").split(",")]
opt = int(os.environ.get("AB_OPT", str(_lib.OPT_LN_FUSE)))
for _ in range(3):
    pipe(rgb, mot)
res = {m: [] for m in modes}
for rnd in range(int(os.environ.get("AB_ROUNDS", "4"))):
    for m in modes:
        ops.set_option(opt, m)
        pipe(rgb, mot)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pipe(rgb, mot)
        e1.record()
        torch.cuda.synchronize()
        res[m].append(e0.elapsed_time(e1) / 3)
for m in modes:
    print(f"option {opt} = {m}: " + " ".join(f"{t:.1f}" for t in res[m]) + f"  | median {sorted(res[m])[len(res[m]) // 2]:.1f} ms/step")
