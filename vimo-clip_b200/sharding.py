"""Clip sharding across the GPUs of one box and the final gather.

The path is embarrassingly parallel over clips: rank r owns clips r, r+W, ... and no collective runs
while encoding.  The only exchange is one ``all_gather_into_tensor`` of the per-rank embeddings /
logits at the end (NCCL over NVLink on GPUs, gloo in the CPU tests), after which the rank-major
concatenation is put back in global clip order with ``indexing.unshard_order``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import indexing


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def local_clip_ids(num_clips: int):
    rank, w = world()
    return indexing.shard_ids(num_clips, rank, w)


def gather_clips(local: torch.Tensor, num_clips: int) -> torch.Tensor:
    """local [n_local, ...] (this rank's clips in ownership order) -> [num_clips, ...] in global order
    on every rank.  Ragged last shards are zero-padded to a common size and trimmed after the gather."""
    rank, w = world()
    if w == 1:
        return local
    per = indexing.padded_per_rank(num_clips, w)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((w * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    order = torch.from_numpy(indexing.unshard_order(num_clips, w)).to(local.device)
    return out.index_select(0, order)
