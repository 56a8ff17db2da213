"""Oracle restatement of the hot-path functions the reference repository itself owns.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED: every function here is checked
against golden vectors produced by executing the reference's own source
(tests/golden/make_golden.py lifts the function bodies out of /root/reference by AST,
without importing the package) and, for ``remove_border_instances``, against the
known-answer cases in the reference's tests/test_remove_border_instances.py:30-117.

  compute_class_masks      <- /root/reference/src/classpose/models.py:191-230
  remove_border_instances  <- /root/reference/src/classpose/metrics/pq.py:65-92
  unaugment_class_tiles    <- /root/reference/src/classpose/transforms/transforms.py:4-21
"""
from __future__ import annotations

import numpy as np


def compute_class_masks(masks: np.ndarray, y_class: np.ndarray):
    """Majority vote of the per-pixel arg-max class inside every instance.

    masks: int label image (any shape); y_class: logits with the class axis first after
    squeezing singleton axes.  Ties go to the lowest class index at both levels; class 0
    (background) may win; label 0 always maps to class 0.  Returns (class_masks int64,
    sorted unique label values of `masks`).
    """
    logits = np.squeeze(y_class)
    n_classes = int(logits.shape[0])
    pix_class = np.argmax(logits, axis=0).reshape(-1)
    inst = masks.reshape(-1)
    top = int(inst.max())
    sel = inst > 0
    table = np.zeros((top + 1, n_classes), dtype=np.int64)
    np.add.at(table, (inst[sel].astype(np.int64), pix_class[sel]), 1)
    winner = np.argmax(table, axis=1)
    winner[0] = 0
    return winner[masks], np.unique(masks)


def remove_border_instances(mask: np.ndarray) -> np.ndarray:
    """Zero (in place) every instance that owns a pixel on the first/last row/column.
    For (H, W, C) input channel 0 holds the instance ids and all channels are zeroed."""
    inst = mask[..., 0] if mask.ndim == 3 else mask
    edge = np.concatenate([inst[0], inst[-1], inst[:, 0], inst[:, -1]])
    ids = np.unique(edge)
    ids = ids[ids != 0]
    mask[np.isin(inst, ids)] = 0
    return mask


def unaugment_class_tiles(y):
    """Undo the parity-pattern flips on class-logit tiles [ny, nx, C, ly, lx]
    (flip only -- no sign change, unlike the flow channels).  numpy-array version."""
    for j in range(y.shape[0]):
        for i in range(y.shape[1]):
            if j % 2 == 0 and i % 2 == 1:
                y[j, i] = y[j, i, :, ::-1, :].copy()
            elif j % 2 == 1 and i % 2 == 0:
                y[j, i] = y[j, i, :, :, ::-1].copy()
            elif j % 2 == 1 and i % 2 == 1:
                y[j, i] = y[j, i, :, ::-1, ::-1].copy()
    return y
