"""Oracle restatement of ``cellpose.utils.fill_holes_and_remove_small_masks`` and the
``fastremap`` / ``fill_voids`` helpers it leans on (cellpose==4.0.8,
fastremap==1.17.7, fill-voids==2.1.1; none on disk -> PARITY UNPINNED).

TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md Appendix A.6.
Reference call sites: /root/reference/src/classpose/models.py:149 (via
resize_and_compute_masks) and :172-174.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import binary_fill_holes, find_objects


def renumber(labels: np.ndarray) -> np.ndarray:
    """fastremap.renumber(in_place=True, preserve_zero=True): contiguous ids 1..n in
    order of first appearance in C (raster) order; 0 stays 0."""
    flat = labels.ravel()
    uniq, first = np.unique(flat, return_index=True)
    keep = uniq != 0
    uniq, first = uniq[keep], first[keep]
    order = np.argsort(first, kind="stable")
    lut = np.zeros(int(flat.max()) + 1 if flat.size else 1, dtype=labels.dtype)
    lut[uniq[order]] = np.arange(1, len(uniq) + 1, dtype=labels.dtype)
    return lut[flat].reshape(labels.shape)


def mask_labels(labels: np.ndarray, to_zero) -> np.ndarray:
    """fastremap.mask: set every listed label value to 0."""
    out = labels.copy()
    out[np.isin(out, np.asarray(to_zero))] = 0
    return out


def fill_voids_2d(m: np.ndarray) -> np.ndarray:
    """fill_voids.fill on a 2-D boolean image: background components that are not
    4-connected to the image border become foreground."""
    return binary_fill_holes(m)


def _drop_small(masks: np.ndarray, min_size: int) -> np.ndarray:
    # counts of the sorted unique values, first entry (assumed to be label 0) dropped;
    # the *position* of a small count, +1, is used as the label value to remove.  That
    # equals the label itself only while labels are contiguous 1..n and 0 is present --
    # an upstream quirk the oracle keeps.
    counts = np.unique(masks, return_counts=True)[1][1:]
    masks = mask_labels(masks, np.nonzero(counts < min_size)[0] + 1)
    return renumber(masks)


def fill_holes_and_remove_small_masks(masks: np.ndarray, min_size: int = 15) -> np.ndarray:
    if masks.ndim != 2:
        raise ValueError("oracle covers 2-D label images only")
    masks = masks.copy()
    if min_size > 0:
        masks = _drop_small(masks, min_size)
    slices = find_objects(masks)
    j = 0
    for i, slc in enumerate(slices):
        if slc is not None:
            msk = masks[slc] == (i + 1)
            msk = fill_voids_2d(msk)
            masks[slc][msk] = j + 1
            j += 1
    if min_size > 0:
        masks = _drop_small(masks, min_size)
    return masks
