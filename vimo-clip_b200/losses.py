"""Drop-in for the reference's ``losses.py`` (forward values; inference / evaluation use).

``distillation_loss(..., mode="cosine")`` runs the fused row-dot/norm reduction kernel (E2) on CUDA
tensors (losses.py:27-40); ``mse`` and ``classification_loss`` (losses.py:47-67) are one-line
reductions left to torch -- they are off the per-frame hot path ("BCE = next").
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def distillation_loss(student_embeddings, teacher_embeddings, mode="mse"):
    if mode == "mse":
        return F.mse_loss(student_embeddings, teacher_embeddings)
    if mode == "cosine":
        if student_embeddings.requires_grad or not student_embeddings.is_cuda:
            # training (autograd) or CPU tensors: same arithmetic as losses.py:23-40, in torch
            eps = 1e-5
            sn = student_embeddings.norm(dim=-1).clamp(min=eps)
            tn = teacher_embeddings.norm(dim=-1).clamp(min=eps)
            cos = (student_embeddings * teacher_embeddings).sum(dim=-1) / (sn * tn)
            return (1 - cos.clamp(-1 + eps, 1 - eps)).mean()
        return ops.cosine_distill_loss(student_embeddings, teacher_embeddings)
    raise ValueError(f"Unsupported mode '{mode}'. Choose 'mse' or 'cosine'.")


def classification_loss(predictions, targets, positive_weight=None):
    num_classes = predictions.shape[-1]
    pos_weight = None
    if positive_weight is not None:
        pos_weight = torch.full((num_classes,), positive_weight, device=predictions.device) * targets + 1
    return F.binary_cross_entropy_with_logits(predictions, targets.float(), pos_weight=pos_weight)


def reconstruction_loss(reconstruction, input):
    raise NotImplementedError
