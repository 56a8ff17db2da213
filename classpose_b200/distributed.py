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


def rank_base_offset(local_total: torch.Tensor, group=None) -> int:
    """Exclusive prefix sum over ranks of the per-rank instance totals (one all_gather of an int64)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = local_total.reshape(1).to(torch.int64)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    return int(torch.stack(gathered).reshape(-1)[:rank].sum().item())


def global_label_offsets(counts: torch.Tensor, engine=None, group=None):
    """counts int32 [B] (this rank's tiles, device tensor) -> (offsets int64 [B], local_total, base):
    global id of label l of tile b = offsets[b] + l."""
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
    return offs + base, int(total.item()), base
