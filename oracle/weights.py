"""Seeded weight generators shared by ``make_golden.py`` and the tests (TEST INFRASTRUCTURE).

No pretrained checkpoints exist offline, so parity is established on seeded random weights.
Every LayerNorm weight/bias is perturbed so the affine parameters are exercised, and the
zero-initialised ``residual_mlp.fc2`` (models/student_model.py:25-26) is randomised, otherwise
that branch would be a no-op ("Value distributions").
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


@torch.no_grad()
def randomise_vit_(vit: nn.Module, seed: int) -> None:
    """In-place seeded init of an OpenAI-layout VisionTransformer.

    Standard deviations follow CLIP's ``initialize_parameters`` so activations stay in a sane
    range through 12-24 pre-LN blocks.
    """
    g = _gen(1000 + seed)
    width = vit.conv1.weight.shape[0]
    layers = len(vit.transformer.resblocks)
    attn_std = width**-0.5
    proj_std = (width**-0.5) * ((2 * layers) ** -0.5)
    fc_std = (2 * width) ** -0.5

    def normal_(t, std):
        t.copy_(torch.randn(t.shape, generator=g) * std)

    normal_(vit.conv1.weight, (3 * vit.conv1.kernel_size[0] ** 2) ** -0.5)
    normal_(vit.class_embedding, width**-0.5)
    normal_(vit.positional_embedding, 0.1)
    normal_(vit.proj, width**-0.5)
    for ln in [vit.ln_pre, vit.ln_post]:
        ln.weight.copy_(1.0 + 0.1 * torch.randn(ln.weight.shape, generator=g))
        ln.bias.copy_(0.1 * torch.randn(ln.bias.shape, generator=g))
    for blk in vit.transformer.resblocks:
        normal_(blk.attn.in_proj_weight, attn_std)
        normal_(blk.attn.in_proj_bias, 0.02)
        normal_(blk.attn.out_proj.weight, proj_std)
        normal_(blk.attn.out_proj.bias, 0.02)
        normal_(blk.mlp.c_fc.weight, fc_std)
        normal_(blk.mlp.c_fc.bias, 0.02)
        normal_(blk.mlp.c_proj.weight, proj_std)
        normal_(blk.mlp.c_proj.bias, 0.02)
        for ln in [blk.ln_1, blk.ln_2]:
            ln.weight.copy_(1.0 + 0.1 * torch.randn(ln.weight.shape, generator=g))
            ln.bias.copy_(0.1 * torch.randn(ln.bias.shape, generator=g))


@torch.no_grad()
def randomise_heads_(student: nn.Module, seed: int) -> None:
    """Seeded init of ``residual_mlp`` and ``classification_head`` (fc2 made non-zero)."""
    g = _gen(2000 + seed)
    for lin in [student.residual_mlp.fc1, student.residual_mlp.fc2, student.classification_head[0], student.classification_head[2]]:
        fan_in = lin.weight.shape[1]
        lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * fan_in**-0.5)
        lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.05)


@torch.no_grad()
def randomise_tfam_(model: nn.Module, seed: int) -> None:
    """Seeded init of every parameter of an ``AMO_CLIP`` (reference or restated: same key names)."""
    g = _gen(3000 + seed)
    for name, p in sorted(model.state_dict().items()):
        if name.endswith("in_proj_weight") or name.endswith(".weight") and p.dim() == 2:
            p.copy_(torch.randn(p.shape, generator=g) * p.shape[1] ** -0.5)
        elif p.dim() == 1 and (".norm_" in name or name == "classifier.0.weight") and name.endswith("weight"):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        elif p.dim() == 1:
            p.copy_(0.05 * torch.randn(p.shape, generator=g))


def state_checksum(sd) -> float:
    """Order-independent float64 checksum used to assert test-time weights == golden-time weights."""
    tot = 0.0
    for k in sorted(sd):
        t = sd[k].detach().double()
        tot += float(t.sum()) + 0.5 * float((t * t).sum())
    return tot
