"""Drop-in for `cellpose.utils.fill_holes_and_remove_small_masks` (reached from
/root/reference/src/classpose/models.py:149 via resize_and_compute_masks, and :172-174)."""
from __future__ import annotations

import numpy as np

from ._abi import ClassposeB200Error
from .engine import get_engine


def fill_holes_and_remove_small_masks(masks, min_size=15, device=None):
    m = np.asarray(masks)
    if m.ndim != 2:
        raise ValueError("classpose_b200 covers 2-D label images only (masks_to_flows 3-D / stitching are out of scope)")
    if m.size == 0 or m.max() <= 0:
        return masks
    eng = get_engine(device)
    out, counts = eng.fill_holes_and_remove_small_masks(np.ascontiguousarray(m.astype(np.int32))[None], int(m.max()) + 2,
                                                        min_size)
    if int(counts[0].item()) < 0:
        raise ClassposeB200Error("hole fill exhausted its bitmap pool (counts = -1): masks would be incomplete")
    return out[0].cpu().numpy().astype(m.dtype)
