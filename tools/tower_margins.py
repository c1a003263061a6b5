"""Embedding cosine of each tower variant against the fp32 oracle (the margin under the 0.9995 bar), a few frames per tower.

    python tools/tower_margins.py      (GPU box; oracle on the host: ~1 minute)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402
from oracle import clip_shim, prologue  # noqa: E402
from vimoclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for name, nfr in (("ViT-B/32", 16), ("ViT-B/16", 8), ("ViT-L/14", 3)):
    o = clip_shim.build_visual(name, seed=11)
    tower = vmc.VisionTower.from_name(name).to(dev)
    tower.load_state_dict(o.state_dict())
    gen = torch.Generator().manual_seed(12)
    u8 = torch.randint(0, 256, (nfr, 3, 224, 224), dtype=torch.uint8, generator=gen)
    with torch.no_grad():
        ref = o(torch.from_numpy(prologue.normalise_u8(u8.numpy())))
    patches = ops.prologue(u8.to(dev), wrap=False, dst="patch", patch=tower.patch_size)
    line = []
    for ln_mode, cls in ((4, 2), (3, 2), (6, 2), (6, 1)):
        tower.ln_mode, tower.last_block_cls = ln_mode, cls
        e = tower.forward_patches(patches, nfr).cpu()
        cos = torch.nn.functional.cosine_similarity(e.double(), ref.double(), dim=-1).min().item()
        rel = ((e - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item()
        line.append(f"ln_mode {ln_mode} cls {cls}: 1-cos {1 - cos:.2e} rel {rel:.2e}")
    print(f"{name} ({nfr} frames): " + " | ".join(line), flush=True)
