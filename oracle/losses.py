"""Restatement of losses.py:5-67 (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def distillation_loss(student, teacher, mode="mse"):
    if mode == "mse":
        return F.mse_loss(student, teacher)
    if mode == "cosine":  # losses.py:23-40
        eps = 1e-5
        sn = student.norm(dim=-1).clamp(min=eps)
        tn = teacher.norm(dim=-1).clamp(min=eps)
        cos = (student * teacher).sum(dim=-1) / (sn * tn)
        cos = cos.clamp(-1 + eps, 1 - eps)
        return (1 - cos).mean()
    raise ValueError(f"Unsupported mode '{mode}'. Choose 'mse' or 'cosine'.")


def classification_loss(predictions, targets, positive_weight=None):  # losses.py:47-67
    c = predictions.shape[-1]
    pw = None
    if positive_weight is not None:
        pw = torch.full((c,), positive_weight, device=predictions.device) * targets + 1
    return F.binary_cross_entropy_with_logits(predictions, targets.float(), pos_weight=pw)
