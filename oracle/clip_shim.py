"""Restatement of the parts of OpenAI ``clip`` the reference uses (TEST INFRASTRUCTURE).

The reference calls ``clip.load(name, device)`` and keeps ``model.visual`` and ``preprocess``
(``models/student_model.py:44-49``).  The package (openai/CLIP @
dcba3cb2e2827b402d2701e7e1c7d9fed8a20ef1, ``requirements.txt:5``) is not installed and there is
no network, so its published architecture is restated here with seeded
random weights.  ``install()`` registers this module as ``clip`` in ``sys.modules`` so the
reference's student files import and run unmodified.

Cross-check: ``tests/test_oracle_golden.py`` maps these weights into the installed HF
``CLIPVisionModelWithProjection`` and requires agreement to ~1e-5, so two independent
implementations pin the restatement.
"""
from __future__ import annotations

import sys
import types
from collections import OrderedDict

import torch
import torch.nn as nn

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

# name -> (patch, width, layers, heads, output_dim); input resolution 224 for all three
VIT_CONFIGS = {
    "ViT-B/32": (32, 768, 12, 12, 512),
    "ViT-B/16": (16, 768, 12, 12, 512),
    "ViT-L/14": (14, 1024, 24, 16, 768),
    # tiny geometry for fast CPU tests (not an OpenAI model)
    "ViT-tiny/32": (32, 128, 2, 2, 64),
}


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    """Pre-LN block: x + attn(ln_1(x)); x + mlp(ln_2(x)).  Sequence-first MultiheadAttention."""

    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(
            OrderedDict(
                [
                    ("c_fc", nn.Linear(d_model, d_model * 4)),
                    ("gelu", QuickGELU()),
                    ("c_proj", nn.Linear(d_model * 4, d_model)),
                ]
            )
        )
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x):
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width**-0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor):
        x = self.conv1(x)  # [F, width, g, g]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # [F, g*g, width]
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)  # sequence first
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


def build_visual(name: str, seed: int = 0) -> VisionTransformer:
    """Seeded random-weight tower (no pretrained weights exist offline)."""
    from .weights import randomise_vit_  # local import: weights imports this module

    patch, width, layers, heads, out = VIT_CONFIGS[name]
    vit = VisionTransformer(224, patch, width, layers, heads, out)
    randomise_vit_(vit, seed)
    return vit.eval()


def _transform(n_px: int):
    """clip/clip.py::_transform: Resize(bicubic) -> CenterCrop -> RGB -> ToTensor -> Normalize."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor

    return Compose(
        [
            Resize(n_px, interpolation=InterpolationMode.BICUBIC),
            CenterCrop(n_px),
            lambda image: image.convert("RGB"),
            ToTensor(),
            Normalize(CLIP_MEAN, CLIP_STD),
        ]
    )


class _ClipModel(nn.Module):
    def __init__(self, visual):
        super().__init__()
        self.visual = visual


_SEED = 0


def set_seed(seed: int) -> None:
    global _SEED
    _SEED = seed


def load(name: str, device="cpu", jit: bool = False, download_root=None):
    """Shim of ``clip.load``: returns ``(model_with_.visual, preprocess)``."""
    model = _ClipModel(build_visual(name, _SEED)).to(device)
    return model, _transform(224)


def available_models():
    return [k for k in VIT_CONFIGS if not k.startswith("ViT-tiny")]


def install() -> types.ModuleType:
    """Register this module as ``clip`` so ``import clip`` in the reference files resolves here."""
    mod = sys.modules[__name__]
    sys.modules["clip"] = mod
    return mod
