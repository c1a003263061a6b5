"""Host-side integer index / mask arithmetic of the path (bit-exact with the reference).

* ``sample_frame_indices``    extract_embeddings.py:77-81
* ``sparse_sampling``         TFAM/data/dataset.py:7-12
* ``collate_fn_pad``          TFAM/data/dataset.py:76-112 (zero pad + boolean masks, True = real frame)
* ``segment_frame_indices``   dataset.py:49-57,80-91 (fixed-length segments, last one padded by repetition)
* ``shard_range`` / ``shard_ids``  clip sharding across ranks
"""
from __future__ import annotations

import numpy as np
import torch


def sample_frame_indices(total_frames: int, max_frames=None) -> np.ndarray:
    if (max_frames is None) or (total_frames <= max_frames):
        return np.arange(total_frames)
    step = total_frames // max_frames
    return np.arange(0, total_frames, step)[:max_frames]


def sparse_sampling(embeddings: torch.Tensor, num_frames: int) -> torch.Tensor:
    total_frames = embeddings.shape[0]
    if total_frames > num_frames:
        # torch.linspace itself, for its float32 rounding
        indices = torch.linspace(0, total_frames - 1, num_frames).long()
        embeddings = embeddings[indices]
    return embeddings


def collate_fn_pad(batch):
    embeddings = [item["embeddings"] for item in batch]
    flow_embeddings = [item["flow_embeddings"] for item in batch]
    labels = torch.stack([item["labels"] for item in batch])
    padded_rgb = torch.nn.utils.rnn.pad_sequence(embeddings, batch_first=True)
    padded_flow = torch.nn.utils.rnn.pad_sequence(flow_embeddings, batch_first=True)
    lens_rgb = torch.tensor([x.shape[0] for x in embeddings])
    lens_flow = torch.tensor([x.shape[0] for x in flow_embeddings])
    mask_rgb = torch.arange(padded_rgb.size(1)).expand(len(lens_rgb), padded_rgb.size(1)) < lens_rgb.unsqueeze(1)
    mask_flow = torch.arange(padded_flow.size(1)).expand(len(lens_flow), padded_flow.size(1)) < lens_flow.unsqueeze(1)
    return {
        "video_id": [item["video_id"] for item in batch],
        "embeddings": padded_rgb,
        "flow_embeddings": padded_flow,
        "labels": labels,
        "mask_rgb": mask_rgb,
        "mask_flow": mask_flow,
    }


def segment_frame_indices(total_frames: int, sequence_length: int):
    segs = []
    for start in range(0, total_frames, sequence_length):
        idx = list(range(start, min(start + sequence_length, total_frames)))
        idx.extend([idx[-1]] * (sequence_length - len(idx)))
        segs.append(idx)
    return segs


def shard_ids(num_clips: int, rank: int, world: int) -> np.ndarray:
    """Round-robin clip ownership r::W."""
    return np.arange(rank, num_clips, world)


def padded_per_rank(num_clips: int, world: int) -> int:
    return (num_clips + world - 1) // world


def unshard_order(num_clips: int, world: int) -> np.ndarray:
    """Index that restores global clip order from the rank-major concatenation of padded shards."""
    per = padded_per_rank(num_clips, world)
    pos = np.empty(num_clips, dtype=np.int64)
    for r in range(world):
        ids = shard_ids(num_clips, r, world)
        pos[ids] = r * per + np.arange(len(ids))
    return pos
