"""Stage 1: CLIP image features, the call surface ``extract_embeddings.py`` uses from HF ``CLIPModel``:

    clip_model = CLIPModel.from_pretrained(...).eval().to(device)          # extract_embeddings.py:17
    embeddings = clip_model.get_image_features(pixel_values)               # :94  -> Tensor [T, D]

``CLIPVisionFeatures`` keeps that surface (``get_image_features`` returns a plain tensor like the
pinned transformers 4.53.2) and loads HF-named weights (``vision_model.*``, ``visual_projection``).
``get_image_features_u8`` is the fused entry: raw uint8 frames, normalisation inside the P1 kernel,
so the fp32 ``pixel_values`` tensor of ``CLIPImageProcessor`` (extract_embeddings.py:89-93) never exists.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .vit import VisionTower, load_hf_vision_state_dict

_HF_NAMES = {
    "openai/clip-vit-base-patch16": "ViT-B/16",
    "openai/clip-vit-base-patch32": "ViT-B/32",
    "openai/clip-vit-large-patch14": "ViT-L/14",
}


class CLIPVisionFeatures(nn.Module):
    def __init__(self, name: str = "openai/clip-vit-base-patch16", **tower_kw):
        super().__init__()
        self.visual = VisionTower.from_name(_HF_NAMES.get(name, name), **tower_kw)

    @classmethod
    def from_hf_state_dict(cls, name: str, state_dict: dict, **kw) -> "CLIPVisionFeatures":
        m = cls(name, **kw)
        load_hf_vision_state_dict(m.visual, state_dict)
        return m

    @torch.no_grad()
    def get_image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """pixel_values fp32 [F,3,224,224] (already normalised) -> [F, D] fp32, NOT L2-normalised."""
        return self.visual(pixel_values)

    @torch.no_grad()
    def get_image_features_u8(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """frames uint8 [F,3,224,224] (RGB, as decord/PIL hand them over) -> [F, D] fp32."""
        if frames_u8.dtype != torch.uint8:
            raise TypeError("get_image_features_u8 expects uint8 frames")
        dev = self.visual.proj.device
        if dev.type != "cuda":
            raise _lib.VmcError("CLIPVisionFeatures runs on CUDA only (no CPU fallback)")
        frames_u8 = frames_u8.to(dev, non_blocking=True)
        res = self.visual.input_resolution
        if frames_u8.shape[-2] != res or frames_u8.shape[-1] != res:
            # CLIPImageProcessor: bicubic resize of the shortest edge to 224 + centre crop (PIL resampler, HF's floor crop offset)
            frames_u8 = ops.resize_center_crop(frames_u8, wrap=False, size=res, hf_crop=True)
        patches = ops.prologue(frames_u8, wrap=False, dst="patch", patch=self.visual.patch_size)
        return self.visual.forward_patches(patches, frames_u8.shape[0])

    forward = get_image_features


class _ProcessorOutput(dict):
    """What ``CLIPImageProcessor.__call__`` returns as far as ``extract_embeddings.py:91-94`` uses it:
    ``inputs = processor(images=..., return_tensors="pt"); inputs = inputs.to(device); model.get_image_features(**inputs)``.
    The frames stay uint8 until ``.to(device)``; resize / crop / rescale / normalise then run on the GPU."""

    def __init__(self, frames_u8: torch.Tensor, size: int):
        super().__init__()
        self._u8, self._size = frames_u8, size

    def to(self, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.VmcError("the image processor runs on CUDA only (no CPU fallback)")
        u8 = self._u8.to(dev, non_blocking=True)
        if u8.shape[-2] != self._size or u8.shape[-1] != self._size:
            u8 = ops.resize_center_crop(u8, wrap=False, size=self._size, hf_crop=True)
        self["pixel_values"] = ops.prologue(u8, wrap=False, dst="f32")
        return self

    @property
    def pixel_values(self):
        return self["pixel_values"]


class CLIPImageProcessor:
    """Drop-in for HF ``CLIPImageProcessor`` on the call surface of ``extract_embeddings.py:18,89-93`` (defaults: bicubic
    resize of the shortest edge to 224, centre crop 224, rescale 1/255, CLIP mean / std): a list of PIL images or
    HxWx3 uint8 arrays of one size -> ``pixel_values`` fp32 [T,3,224,224], computed by the Pillow-exact resize kernel and
    the P1 prologue on the GPU ((u8/255 - mean)/std with true fp32 divisions; within 1 ulp of the HF slow path)."""

    def __init__(self, size: int = 224):
        self.size = size

    @classmethod
    def from_pretrained(cls, *_args, **_kw):
        return cls()

    def __call__(self, images, return_tensors: str = "pt"):
        import numpy as np

        if return_tensors != "pt":
            raise ValueError("only return_tensors='pt' is supported")
        if not isinstance(images, (list, tuple)):
            images = [images]
        arrs = [np.asarray(im.convert("RGB") if hasattr(im, "convert") else im) for im in images]
        if any(a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3 for a in arrs) or len({a.shape for a in arrs}) != 1:
            raise ValueError("expected RGB uint8 images (PIL or HxWx3 arrays) of one common size")
        frames = torch.from_numpy(np.stack(arrs)).permute(0, 3, 1, 2).contiguous()
        return _ProcessorOutput(frames, self.size)
