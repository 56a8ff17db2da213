"""Next row N2: drop-in for `deduplicate` of the WSI entry point
(/root/reference/src/classpose/entrypoints/predict_wsi.py:896-965) -- cells detected twice where tiles overlap are
reduced to the largest one.  The pair search and grouping run on the device (grid hash + union-find)."""
from __future__ import annotations

import numpy as np
import torch

from .engine import get_engine


def keep_mask(centers, sizes, max_dist: float = 15 / 2, device=None):
    """centers [n,2] (x, y), sizes [n] -> boolean numpy keep mask (device tensors are accepted)."""
    eng = get_engine(device)
    c = centers if isinstance(centers, torch.Tensor) else torch.as_tensor(np.asarray(centers, np.float64))
    s = sizes if isinstance(sizes, torch.Tensor) else torch.as_tensor(np.asarray(sizes, np.float64))
    if c.shape[0] == 0:
        return np.zeros(0, bool)
    c = c.to(eng.device, torch.float64)
    keep, _ = eng.dedup_cells(c[:, 0].contiguous(), c[:, 1].contiguous(), s, max_dist)
    return keep.cpu().numpy().astype(bool)


def deduplicate(features: list[dict], max_dist: float = 15 / 2, device=None) -> list[dict]:
    """Same signature and feature layout as the reference: every feature carries
    feature["properties"]["measurements"] = [{"name": "area" | "centroidX" | "centroidY", "value": ...}, ...]."""
    if not features:
        return []
    centers, sizes = [], []
    for feature in features:
        m = {x["name"]: x["value"] for x in feature["properties"]["measurements"]}
        sizes.append(m["area"])
        centers.append([m["centroidX"], m["centroidY"]])
    keep = keep_mask(np.asarray(centers, np.float64), np.asarray(sizes, np.float64), max_dist, device)
    return [f for f, k in zip(features, keep) if k]
