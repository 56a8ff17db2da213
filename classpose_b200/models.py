"""Drop-ins for the two module-level functions of `classpose.models` on the hot path
(hooks A and C of SURVEY.md 8b): same signatures, argument meaning and return types.

  compute_masks        <- /root/reference/src/classpose/models.py:97-188
  compute_class_masks  <- /root/reference/src/classpose/models.py:191-230
"""
from __future__ import annotations

import numpy as np
import torch

from . import dynamics
from .engine import get_engine


def compute_masks(dP, cellprob, shape, do_3D, niter, cellprob_threshold, flow_threshold, min_size,
                  max_size_fraction, stitch_threshold, device):
    """2-D branch of classpose.models.compute_masks: every plane of dP [2,nimg,H,W] / cellprob [nimg,H,W]
    is one independent tile; all planes run as one batch.  Returns the single [H,W] array when nimg == 1."""
    if do_3D:
        raise NotImplementedError("do_3D=True is outside the B200 hot path (2-D tiles only)")
    nimg = int(shape[0])
    if stitch_threshold > 0 and nimg > 1:
        raise NotImplementedError("3-D stitching (stitch_threshold > 0) is outside the B200 hot path")
    dP = np.asarray(dP, np.float32)
    cellprob = np.asarray(cellprob, np.float32)
    if dP.ndim != 4 or cellprob.ndim != 3 or dP.shape[1] != cellprob.shape[0]:
        raise ValueError(f"expected dP [2,nimg,H,W] and cellprob [nimg,H,W], got {dP.shape} and {cellprob.shape}")
    kw = dict(niter=niter, cellprob_threshold=cellprob_threshold, flow_threshold=flow_threshold, min_size=min_size,
              max_size_fraction=max_size_fraction, resize=None, device=device)
    if nimg == 1:
        # the WSI loop's case (predict_wsi.py:750-751): one tile; the array returned here is the very object
        # compute_class_masks receives next (models.py:753-768), which lets the vote reuse the labels on the device
        return dynamics.resize_and_compute_masks(np.ascontiguousarray(dP[:, 0]), np.ascontiguousarray(cellprob[0]), **kw)
    return dynamics.resize_and_compute_masks(np.ascontiguousarray(dP.transpose(1, 0, 2, 3)), cellprob, **kw)


def compute_class_masks(masks, y_class, device=None):
    """Per-instance majority vote of the per-pixel arg-max class.  Ties -> lowest class index, class 0
    may win, label 0 -> class 0.  Returns (class_masks int64, np.unique(masks)).  As in the reference the
    vote table is indexed by label value over the whole array: planes of a stack that reuse an id vote
    together."""
    if isinstance(masks, np.ndarray):
        from . import fastpath
        plan = fastpath.plan_holding(masks)
        lg = np.asarray(y_class)
        if plan is not None and lg.size % masks.size == 0 and lg.size // masks.size >= 1:
            # `masks` is the array hook A / B returned last on this thread: its labels are still on the device, only
            # the logits travel (models.py:753-768 hands the array straight through)
            cm = plan.vote(lg.reshape(lg.size // masks.size, masks.shape[0], masks.shape[1])).astype(np.int64)
            if plan.contiguous_ids:
                n = plan.last_count
                first = 0 if (n == 0 or masks.min() == 0) else 1
                unique_instances = np.arange(first, n + 1, dtype=masks.dtype)
            else:
                unique_instances = np.unique(masks)
            return cm, unique_instances
    masks = np.asarray(masks)
    logits = np.squeeze(np.asarray(y_class, np.float32))
    C = int(logits.shape[0])
    if masks.size != logits[0].size:
        raise ValueError(f"masks {masks.shape} and y_class {np.shape(y_class)} cover different pixel counts")
    unique_instances = np.unique(masks)
    top = int(masks.max()) if masks.size else 0
    if top <= 0:
        return np.zeros(masks.shape, np.int64), unique_instances
    W = masks.shape[-1]
    rows = masks.size // W
    eng = get_engine(device)
    m = np.ascontiguousarray(masks.astype(np.int32)).reshape(1, rows, W)
    g = np.ascontiguousarray(logits).reshape(1, C, rows, W)
    _, cm = eng.class_vote(m, g, top + 2, want_class_masks=True)
    return cm.cpu().numpy().reshape(masks.shape).astype(np.int64), unique_instances
