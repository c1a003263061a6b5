"""MoCLIP student encoders: drop-ins for ``models/student_model.py`` and
``models/student_model_frame_diff.py`` (constructor, forward signature, attributes and state_dict
keys identical; reference checkpoints with the ``module.`` prefix of ``train.py:167`` load after
stripping the prefix or through ``nn.DataParallel``).

forward(videos [B,T,3,224,224] uint8 | float32) -> (embeddings [B,T,D], embeddings_for_distillation
[B,T,D], logits [B,num_classes]), all fp32 on the model's device:

  1. P1 prologue kernel: the reference's per-frame ``to_pil_image`` wrap + CLIP normalise
     (student_model.py:74-81) done once on the GPU, emitted directly as the bf16 patch matrix;
  2. CLIP ViT tower on tcgen05 GEMM / attention kernels (student_model.py:84);
  3. heads: ResidualMLP ``x + alpha*fc2(GELU(fc1(x)))`` (:33-35,90), temporal mean of the RAW
     embeddings (:93) and the 2-layer classification head (:96) as GEMM epilogues.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from ._params import SharedCache
from .vit import VisionTower


class ResidualMLP(nn.Module):
    """Parameters of models/student_model.py:8-35 (fc2 zero-initialised, fixed scalar alpha)."""

    def __init__(self, embed_dim: int, alpha: float = 0.1):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, embed_dim)
        self.fc2 = nn.Linear(embed_dim, embed_dim)
        self.alpha = alpha
        nn.init.zeros_(self.fc2.weight)
        nn.init.zeros_(self.fc2.bias)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).float().contiguous()
        h = ops.gemm(ops.cast_bf16(x2, split=True), ops.split_weight(self.fc1.weight), bias=self.fc1.bias.float(),
                     act=ops.ACT_GELU_ERF, out_dtype=torch.float32)
        y = ops.gemm(ops.cast_bf16(h, split=True), ops.split_weight(self.fc2.weight), bias=self.fc2.bias.float(),
                     alpha=float(self.alpha), resid=x2, out_dtype=torch.float32)
        return y.view(shp)


class _Preprocess:
    """``clip.load``'s second return value as seen by callers (inference.py:90 reads ``.transforms``)."""

    def __init__(self, n_px: int = 224):
        self._n_px = n_px
        self._compose = None

    def _build(self):
        if self._compose is None:
            from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor

            self._compose = Compose([
                Resize(self._n_px, interpolation=InterpolationMode.BICUBIC),
                CenterCrop(self._n_px),
                lambda image: image.convert("RGB"),
                ToTensor(),
                Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
            ])
        return self._compose

    @property
    def transforms(self):
        return self._build().transforms

    def __call__(self, img):
        return self._build()(img)


class _StudentBase(nn.Module):
    def __init__(self, clip_model_name: str = "ViT-B/32", device="cuda", num_classes: int = 140, alpha: float = 0.1):
        super().__init__()
        self.device = device
        # The reference gets pretrained weights from clip.load(); offline there are none, so the tower
        # starts from a seeded-style random init and takes weights through load_state_dict().
        self.preprocess = _Preprocess(224)
        self.visual_encoder = VisionTower.from_name(clip_model_name)
        embed_dim = self.visual_encoder.output_dim
        self.residual_mlp = ResidualMLP(embed_dim, alpha=alpha)
        self.classification_head = nn.Sequential(
            nn.Linear(embed_dim, embed_dim // 2),
            nn.ReLU(),
            nn.Linear(embed_dim // 2, num_classes),
        )
        self._head_cache = SharedCache(self)
        self.to(device)

    def _heads(self):
        ps = [self.residual_mlp.fc1.weight, self.residual_mlp.fc1.bias, self.residual_mlp.fc2.weight, self.residual_mlp.fc2.bias,
              self.classification_head[0].weight, self.classification_head[0].bias,
              self.classification_head[2].weight, self.classification_head[2].bias]

        def build():
            # one fused fp32 kernel per clip (vmc_student_heads): weights transposed to [K, N] fp32 once
            return [p.detach().float().t().contiguous() if p.dim() == 2 else p.detach().float().contiguous() for p in ps]

        self._head_cache.bind(self)
        return self._head_cache.get(("heads", ps[0].device.index), self._head_cache.signature(self), build)

    @torch.no_grad()
    def encode_patches(self, patches: torch.Tensor, B: int, T: int):
        """bf16 patch matrix of B*T frames -> the three outputs of the reference forward."""
        emb = self.visual_encoder.forward_patches(patches, B * T)  # [B*T, D] fp32
        D = emb.shape[1]
        num_classes = self.classification_head[2].weight.shape[0]
        distill, logits = ops.student_heads(emb.view(B, T, D), self._heads(), float(self.residual_mlp.alpha), num_classes)
        return emb.view(B, T, D), distill.view(B, T, D), logits

    def forward(self, videos: torch.Tensor):
        """``.eval()``: inference kernels under no_grad.  ``.train()`` (train.py:86-107): the three outputs carry a backward
        through our kernels for every parameter of the tower and the heads (``student_train.py``)."""
        train = self.training and torch.is_grad_enabled()
        with torch.no_grad():
            patches, B, T = self._patches(videos)
            if not train:
                return self.encode_patches(patches, B, T)
        from .student_train import student_train_forward

        return student_train_forward(self, patches, B, T)

    def _patches(self, videos: torch.Tensor):
        if videos.dim() != 5 or videos.shape[2] != 3:
            raise ValueError("expected videos of shape (B, T, 3, H, W)")
        B, T, C, H, W = videos.shape
        res = self.visual_encoder.input_resolution
        dev = self.visual_encoder.proj.device
        if dev.type != "cuda":
            raise _lib.VmcError("the student runs on CUDA only (no CPU fallback)")
        frames = videos.reshape(B * T, C, H, W)
        if frames.dtype != torch.uint8:
            frames = frames.float()  # student_model.py:74
        frames = frames.to(dev, non_blocking=True)
        if H != res or W != res:
            # Resize(224, BICUBIC) -> CenterCrop(224) of the reference preprocess, after the to_pil_image wrap; smaller frames are
            # scaled UP by the same resampler (the short side becomes 224, so CenterCrop never pads)
            u8 = ops.resize_center_crop(frames, wrap=True, size=res)
            patches = ops.prologue(u8, wrap=False, dst="patch", patch=self.visual_encoder.patch_size)
        else:
            patches = ops.prologue(frames, wrap=True, dst="patch", patch=self.visual_encoder.patch_size)
        return patches, B, T


class FlowStudentModel(_StudentBase):
    """Drop-in for ``models.student_model.FlowStudentModel`` (models/student_model.py:38-98)."""


class FrameDiffStudentModel(_StudentBase):
    """Drop-in for ``models.student_model_frame_diff.FrameDiffStudentModel`` (:38-86).

    Adds ``forward_bgr``: raw BGR clips -> frame difference (utils/generate_frame_diff_video.py:37-49)
    -> student, fused in one prologue kernel (the reference round-trips through an H.264 file).
    """

    @torch.no_grad()
    def forward_bgr(self, bgr: torch.Tensor):
        """bgr [B, T+1, 224, 224, 3] uint8 -> outputs for the T difference frames of each clip."""
        dev = self.visual_encoder.proj.device
        bgr = bgr.to(dev, non_blocking=True)
        B, T1 = bgr.shape[0], bgr.shape[1]
        _, patches = ops.frame_diff(bgr, dst="patch", patch=self.visual_encoder.patch_size, want_diff=False)
        return self.encode_patches(patches, B, T1 - 1)
