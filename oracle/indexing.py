"""Integer index / mask arithmetic of the path, restated (TEST INFRASTRUCTURE; bit-exact).

* ``sample_frame_indices``  -- extract_embeddings.py:77-81
* ``sparse_sampling_indices`` -- TFAM/data/dataset.py:7-12 (``torch.linspace(0, T-1, n).long()``)
* ``pad_and_mask``          -- TFAM/data/dataset.py:76-112 (``collate_fn_pad``)
* ``segment_indices``       -- dataset.py:49-57,80-91 (segments of ``sequence_length``, pad by repeating last)
* ``shard_clips``           -- clip sharding r::W of the multi-GPU path (new)
"""
from __future__ import annotations

import numpy as np
import torch


def sample_frame_indices(total_frames: int, max_frames):
    if max_frames is None or total_frames <= max_frames:
        return np.arange(total_frames)
    step = total_frames // max_frames
    return np.arange(0, total_frames, step)[:max_frames]


def sparse_sampling_indices(total_frames: int, num_frames: int) -> torch.Tensor:
    if total_frames > num_frames:
        return torch.linspace(0, total_frames - 1, num_frames).long()
    return torch.arange(total_frames)


def pad_and_mask(seqs):
    """list of [T_i, D] -> (padded [B, T_max, D], mask [B, T_max] bool, True = real frame)."""
    lens = torch.tensor([s.shape[0] for s in seqs])
    t_max = int(lens.max())
    padded = torch.zeros(len(seqs), t_max, seqs[0].shape[1], dtype=seqs[0].dtype)
    for i, s in enumerate(seqs):
        padded[i, : s.shape[0]] = s
    mask = torch.arange(t_max).expand(len(seqs), t_max) < lens.unsqueeze(1)
    return padded, mask


def segment_indices(total_frames: int, sequence_length: int):
    """Frame indices of each segment; the last one is padded by repeating the final frame."""
    out = []
    for start in range(0, total_frames, sequence_length):
        idx = list(range(start, min(start + sequence_length, total_frames)))
        idx += [idx[-1]] * (sequence_length - len(idx))
        out.append(idx)
    return out


def shard_clips(num_clips: int, rank: int, world: int):
    """Clip ids owned by ``rank`` (round-robin r::W) and the padded per-rank count."""
    ids = np.arange(rank, num_clips, world)
    per_rank = (num_clips + world - 1) // world
    return ids, per_rank
