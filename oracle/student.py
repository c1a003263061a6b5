"""fp32 CPU restatement of the MoCLIP student forward (TEST INFRASTRUCTURE).

Follows ``models/student_model.py:8-98`` (``ResidualMLP``, ``FlowStudentModel``) and the identical
``models/student_model_frame_diff.py:8-86`` (``FrameDiffStudentModel``).  The per-frame PIL loop of
``student_model.py:77-78`` is restated arithmetically in ``oracle.prologue`` (wrap -> /255 ->
normalise); at 224x224 the Resize / CenterCrop are identities.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import clip_shim, prologue


class ResidualMLP(nn.Module):
    """x + alpha * fc2(GELU_erf(fc1(x)))  (models/student_model.py:8-35)."""

    def __init__(self, embed_dim: int, alpha: float = 0.1):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, embed_dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(embed_dim, embed_dim)
        self.alpha = alpha
        nn.init.zeros_(self.fc2.weight)
        nn.init.zeros_(self.fc2.bias)

    def forward(self, x):
        return x + self.alpha * self.fc2(self.act(self.fc1(x)))


class StudentOracle(nn.Module):
    """State-dict compatible with the reference ``FlowStudentModel`` / ``FrameDiffStudentModel``."""

    def __init__(self, clip_model_name: str = "ViT-B/32", num_classes: int = 140, alpha: float = 0.1, seed: int = 0):
        super().__init__()
        self.visual_encoder = clip_shim.build_visual(clip_model_name, seed)
        d = self.visual_encoder.output_dim
        self.residual_mlp = ResidualMLP(d, alpha=alpha)
        self.classification_head = nn.Sequential(nn.Linear(d, d // 2), nn.ReLU(), nn.Linear(d // 2, num_classes))

    def forward(self, videos: torch.Tensor):
        """videos [B,T,3,224,224] uint8 or float -> (emb [B,T,D], emb_distill [B,T,D], logits [B,C]).
        Inference under no_grad; in ``.train()`` mode autograd stays on (train.py:95-107 backpropagates through the
        reference module) -- the oracle of the student training-step parity test."""
        with torch.set_grad_enabled(self.training and torch.is_grad_enabled()):
            return self._forward(videos)

    def _forward(self, videos: torch.Tensor):
        B, T, C, H, W = videos.shape
        frames = videos.reshape(B * T, C, H, W)  # student_model.py:74 (.float() happens in preprocess_frames)
        x = torch.from_numpy(prologue.preprocess_frames(frames.cpu().numpy()))  # :77-78
        emb = self.visual_encoder(x).view(B, T, -1)  # :84,87
        distill = self.residual_mlp(emb)  # :90
        pooled = emb.mean(dim=1)  # :93 pools the RAW embeddings
        logits = self.classification_head(pooled.float())  # :96
        return emb, distill, logits
