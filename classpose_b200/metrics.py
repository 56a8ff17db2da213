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
    nch = mask.shape[2] if mask.ndim == 3 else 1
    eng = get_engine(device)
    m = np.ascontiguousarray(mask.astype(np.int32)).reshape(1, mask.shape[0], mask.shape[1], nch)
    out = eng.remove_border_instances(m, int(inst.max()) + 2, nch=nch)
    mask[...] = out.cpu().numpy().reshape(mask.shape).astype(mask.dtype)
    return mask
