"""Full ViMoCLIP inference chained in HBM (BASELINE config 4): CLIP ViT on RGB frames + MoCLIP student
on motion frames + TFAM fusion -> multi-label logits.  The reference decouples the three stages through
HDF5 files (extract_embeddings.py -> inference.py -> TFAM/train_and_eval.py); here the embeddings never
leave the GPU.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import sharding
from .clip_hf import CLIPVisionFeatures
from .student import FlowStudentModel, FrameDiffStudentModel
from .tfam import AMO_CLIP


class ViMoCLIPPipeline(nn.Module):
    def __init__(self, rgb_model: str = "openai/clip-vit-base-patch16", student_model: str = "ViT-B/32", num_classes: int = 140,
                 frame_diff: bool = False, device="cuda", clips_per_step: int = 256):
        super().__init__()
        self.rgb = CLIPVisionFeatures(rgb_model).to(device)
        cls = FrameDiffStudentModel if frame_diff else FlowStudentModel
        self.student = cls(student_model, device=device, num_classes=num_classes)
        self.tfam = AMO_CLIP(num_classes=num_classes, device=device).to(device).eval()
        self.clips_per_step = clips_per_step
        self.device = device

    @torch.no_grad()
    def forward(self, rgb_u8: torch.Tensor, motion_u8: torch.Tensor, mask_rgb=None, mask_flow=None):
        """rgb_u8 [N,T,3,224,224] uint8 RGB frames; motion_u8 [N,T_m,3,224,224] uint8 flow / frame-diff frames
        (host or device).  Returns (logits [N,C], rgb_emb [N,T,D], motion_emb [N,T_m,D]) on the device.

        Host inputs are copied clip-chunk by clip-chunk (``clips_per_step``) so the H2D copy of chunk i+1
        is queued behind the kernels of chunk i on the same stream; the towers batch every frame of a chunk."""
        N, T = rgb_u8.shape[:2]
        e_rgb, e_mot = [], []
        for c0 in range(0, N, self.clips_per_step):
            r = rgb_u8[c0:c0 + self.clips_per_step].to(self.device, non_blocking=True)
            m = motion_u8[c0:c0 + self.clips_per_step].to(self.device, non_blocking=True)
            n = r.shape[0]
            e_rgb.append(self.rgb.get_image_features_u8(r.reshape(n * T, *r.shape[2:])).view(n, T, -1))
            e_mot.append(self.student(m)[0])
        er = e_rgb[0] if len(e_rgb) == 1 else torch.cat(e_rgb)
        em = e_mot[0] if len(e_mot) == 1 else torch.cat(e_mot)
        logits = self.tfam(er, em, mask_rgb, mask_flow)
        return logits, er, em

    @torch.no_grad()
    def forward_sharded(self, rgb_u8_local, motion_u8_local, num_clips: int):
        """Each rank passes only the clips it owns (``sharding.local_clip_ids``); every rank gets the
        gathered logits / embeddings in global clip order (one NCCL all-gather each)."""
        lg, er, em = self.forward(rgb_u8_local, motion_u8_local)
        return (sharding.gather_clips(lg, num_clips), sharding.gather_clips(er, num_clips), sharding.gather_clips(em, num_clips))
