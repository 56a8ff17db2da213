"""Drop-in for `classpose.metrics.pq.remove_border_instances`
(/root/reference/src/classpose/metrics/pq.py:65-92)."""
from __future__ import annotations

import numpy as np

from .engine import get_engine


def remove_border_instances(mask: np.ndarray, device=None) -> np.ndarray:
    """Zero every instance touching the first/last row/column.  (H, W) or (H, W, C) with the instance
    ids in channel 0; all channels are zeroed.  Mutates and returns `mask`, like the reference."""
    if mask.ndim not in (2, 3):
        raise ValueError("mask must be (H, W) or (H, W, C)")
    inst = mask[..., 0] if mask.ndim == 3 else mask
    if inst.size == 0 or inst.max() <= 0:
        return mask
    # Only the instance channel goes to the device, as int32 ids: directly when the dtype allows it, otherwise
    # (float masks, ids >= 2^31) through the ranks of the distinct values.  The device returns which pixels
    # belong to a border instance; those are zeroed here in the caller's array and dtype, so surviving pixels
    # -- every channel of them -- keep their exact values, as in the reference.
    if np.issubdtype(inst.dtype, np.integer) and int(inst.max()) < 2 ** 31 - 2 and int(inst.min()) >= 0:
        ids = np.ascontiguousarray(inst, dtype=np.int32)
    else:
        uniq, inv = np.unique(inst, return_inverse=True)
        rank = np.arange(1, len(uniq) + 1, dtype=np.int32)      # every distinct value is an instance ...
        rank[uniq == 0] = 0                                     # ... except 0, the background (pq.py:88)
        ids = np.ascontiguousarray(rank[inv].reshape(inst.shape), dtype=np.int32)
    eng = get_engine(device)
    out = eng.remove_border_instances(ids[None], int(ids.max()) + 2, nch=1)
    removed = (ids > 0) & (out[0].cpu().numpy() == 0)
    mask[removed] = 0
    return mask
