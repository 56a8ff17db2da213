"""Instance-matching metric used by the parity harness.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates the definition in
/root/reference/src/classpose/metrics/stats_utils.py:64-178 (``get_pq``): instances
pair iff IoU > match_iou (unique for match_iou >= 0.5) and
F1 = DQ = 2TP / (2TP + FP + FN)  (/root/reference/src/classpose/metrics/pq.py:150).
Computed here from one joint histogram instead of per-instance crops.
"""
from __future__ import annotations

import numpy as np


def _contiguous(lab: np.ndarray):
    vals, inv = np.unique(lab, return_inverse=True)
    inv = inv.reshape(lab.shape)
    if vals.size and vals[0] == 0:
        return inv, len(vals) - 1
    return inv + 1, len(vals)


def match_instances(true: np.ndarray, pred: np.ndarray, match_iou: float = 0.5):
    """Returns dict(tp, fp, fn, f1, pairs) where pairs is a list of (true_value,
    pred_value, iou) in the *original* label values."""
    t_vals = np.unique(true); t_vals = t_vals[t_vals != 0]
    p_vals = np.unique(pred); p_vals = p_vals[p_vals != 0]
    t, nt = _contiguous(true)
    p, npd = _contiguous(pred)
    joint = np.zeros((nt + 1, npd + 1), np.int64)
    np.add.at(joint, (t.ravel(), p.ravel()), 1)
    area_t = joint.sum(1)
    area_p = joint.sum(0)
    inter = joint[1:, 1:].astype(np.float64)
    union = area_t[1:, None] + area_p[None, 1:] - inter
    iou = np.where(union > 0, inter / np.maximum(union, 1), 0.0)
    ti, pi = np.nonzero(iou > match_iou)
    tp = len(ti)
    fn = nt - len(set(ti.tolist()))
    fp = npd - len(set(pi.tolist()))
    denom = 2 * tp + fp + fn
    f1 = 1.0 if denom == 0 else 2 * tp / denom
    pairs = [(int(t_vals[a]), int(p_vals[b]), float(iou[a, b])) for a, b in zip(ti, pi)]
    return dict(tp=tp, fp=fp, fn=fn, f1=f1, pairs=pairs, n_true=nt, n_pred=npd)


def class_agreement(true_masks, true_cls, pred_masks, pred_cls, match_iou=0.5):
    """Class parity on matched cells.  `*_cls` map label value -> class (array indexed by
    label value, or a class-mask image of the same shape as the label image)."""
    m = match_instances(true_masks, pred_masks, match_iou)

    def cls_of(masks, cls, v):
        if cls.shape == masks.shape:
            return int(cls[masks == v][0])
        return int(cls[v])

    bad = [(a, b) for a, b, _ in m["pairs"]
           if cls_of(true_masks, true_cls, a) != cls_of(pred_masks, pred_cls, b)]
    m["class_mismatch"] = bad
    return m
