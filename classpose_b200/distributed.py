"""Multi-GPU layout of the post-network path: tiles are independent, so each rank owns a
contiguous shard of the tile list and nothing crosses GPUs on the data path.  The one exchange
is the per-rank instance total, all-gathered (8 bytes per rank over NCCL/NVLink, or gloo in the
CPU tests) so that every rank can turn its per-tile counts into global label offsets.

The reference has no counterpart (ids are per-tile, cells get uuid4 strings:
/root/reference/src/classpose/entrypoints/predict_wsi.py:644); its workers are one process per
GPU pulling tiles from a shared queue (predict_wsi.py:1542-1572).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_tiles: int, rank: int, world_size: int):
    """Contiguous [start, stop) of the tile list owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_tiles, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_base_offset(local_total: torch.Tensor, group=None) -> torch.Tensor:
    """Exclusive prefix sum over ranks of the per-rank instance totals: one all_gather_into_tensor of an int64 per
    rank.  Returns a 0-d int64 tensor ON THE DEVICE of `local_total` -- nothing is read back to the host, so the
    exchange stays asynchronous on the stream (callers that want a python int call .item() themselves)."""
    mine = local_total.reshape(1).to(torch.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return torch.zeros((), dtype=torch.int64, device=mine.device)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = torch.empty((world,), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return gathered[:rank].sum()


def global_label_offsets(counts: torch.Tensor, engine=None, group=None):
    """counts int32 [B] (this rank's tiles, device tensor) -> (offsets int64 [B], local_total, base), all tensors on
    the device of `counts`: global id of label l of tile b = offsets[b] + l.  No host synchronisation."""
    if counts.is_cuda:
        if engine is None:
            from .engine import get_engine
            engine = get_engine(counts.device)
        offs, total = engine.label_offsets(counts, 0)
    else:  # host-side logic (tests): same arithmetic without a device
        c64 = counts.to(torch.int64)
        offs = torch.cumsum(c64, 0) - c64
        total = c64.sum().reshape(1)
    base = rank_base_offset(total, group)
    return offs + base, total.reshape(()), base
