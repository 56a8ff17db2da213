"""Seeded synthetic network outputs for parity tests (numpy, CPU).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows SURVEY.md section 8(d): plant
non-overlapping ellipses, derive ground-truth flows from them with the oracle's
``masks_to_flows`` (x5 = network scale), cellprob = +-6, logits +4 on the true class,
then add Gaussian noise.  ``adversarial_labels`` builds the edge-case label image
(touching cells, border cells, an over-sized cell, an enclosed cell, tiny cells).
"""
from __future__ import annotations

import numpy as np

from . import dynamics


def plant_ellipses(rng, H, W, n_grid, axes=(5.0, 9.0), drop=0.1):
    """One ellipse per jittered grid cell; each stays inside its grid cell, so cells
    never touch.  Returns int32 label image (labels 1..n in grid raster order)."""
    gy, gx = H / n_grid, W / n_grid
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    lab = np.zeros((H, W), np.int32)
    k = 0
    for j in range(n_grid):
        for i in range(n_grid):
            a, b = rng.uniform(axes[0], axes[1], size=2)
            th = rng.uniform(0, np.pi)
            jit_y = max(gy / 2 - max(a, b) - 1.0, 0.0)
            jit_x = max(gx / 2 - max(a, b) - 1.0, 0.0)
            cy = (j + 0.5) * gy + rng.uniform(-jit_y, jit_y)
            cx = (i + 0.5) * gx + rng.uniform(-jit_x, jit_x)
            keep = rng.uniform() >= drop
            if not keep:
                continue
            dy, dx = yy - cy, xx - cx
            u = (dx * np.cos(th) + dy * np.sin(th)) / a
            v = (-dx * np.sin(th) + dy * np.cos(th)) / b
            k += 1
            lab[(u * u + v * v) <= 1.0] = k
    return lab


def outputs_from_labels(rng, lab, C, sigma_flow=0.5, sigma_prob=1.0, sigma_logit=1.0,
                        cell_class=None):
    """Network-like outputs for a planted label image.
    Returns dP [2,H,W] f32, cellprob [H,W] f32, logits [C,H,W] f32, cell_class [n+1]."""
    H, W = lab.shape
    n = int(lab.max())
    mu = dynamics.masks_to_flows(lab) if n > 0 else np.zeros((2, H, W))
    dP = (5.0 * mu + rng.normal(0, sigma_flow, size=(2, H, W))).astype(np.float32)
    cellprob = (np.where(lab > 0, 6.0, -6.0) + rng.normal(0, sigma_prob, size=(H, W))).astype(np.float32)
    if cell_class is None:
        cell_class = np.zeros(n + 1, np.int64)
        if C > 1:
            cell_class[1:] = rng.integers(1, C, size=n)
    cls_img = cell_class[lab]
    logits = rng.normal(0, sigma_logit, size=(C, H, W))
    np.put_along_axis(logits, cls_img[None], np.take_along_axis(logits, cls_img[None], 0) + 4.0, 0)
    return dP, cellprob, logits.astype(np.float32), cell_class


def make_tile(seed, H=256, W=256, C=7, n_grid=10, axes=(5.0, 9.0), drop=0.1,
              sigma_flow=0.5, sigma_prob=1.0):
    """The standard synthetic tile: ~n_grid^2*(1-drop) nuclei.  conic: C=7, 256^2, n_grid=10,
    axes (5,9); dense stress: 512^2, n_grid=45, axes (3.5,5)."""
    rng = np.random.default_rng(1234 + seed)
    lab = plant_ellipses(rng, H, W, n_grid, axes, drop)
    dP, cellprob, logits, cell_class = outputs_from_labels(rng, lab, C, sigma_flow, sigma_prob)
    return dict(labels=lab, dP=dP, cellprob=cellprob, logits=logits, cell_class=cell_class)


def adversarial_labels(H=256, W=256):
    """Edge-case label image: touching cells, border cells, an over-sized cell (> 40 % of
    the tile), a ring cell enclosing background and another cell, tiny cells (< 15 px)."""
    yy, xx = np.mgrid[0:H, 0:W]
    lab = np.zeros((H, W), np.int32)
    k = 0

    def disk(cy, cx, r):
        return (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r

    # over-sized cell: area > 0.4*H*W when H=W=256 (r=95 -> 28.3k px > 26.2k)
    k += 1; lab[disk(H * 0.62, W * 0.40, 0.371 * min(H, W))] = k
    # ring enclosing a hole with background + an inner cell (overwrites the big cell locally)
    ring = disk(40, W - 50, 30) & ~disk(40, W - 50, 18)
    lab[disk(40, W - 50, 31)] = 0
    k += 1; lab[ring] = k
    k += 1; lab[disk(40, W - 50, 8)] = k
    # two touching half disks
    d = disk(40, 45, 12)
    lab[disk(40, 45, 13)] = 0
    k += 1; lab[d & (xx < 45)] = k
    k += 1; lab[d & (xx >= 45)] = k
    # border cells (top edge, left edge, corner)
    for cy, cx in ((0, 120), (128, 0), (H - 1, W - 1), (H - 1, 90)):
        m = disk(cy, cx, 9)
        lab[disk(cy, cx, 10)] = 0
        k += 1; lab[m] = k
    # tiny cells
    for cy, cx, r in ((20, 100, 1.5), (20, 140, 2.0), (90, W - 20, 1.0)):
        m = disk(cy, cx, r)
        lab[disk(cy, cx, r + 1)] = 0
        k += 1; lab[m] = k
    # cell with an internal pinhole (hole fill target)
    m = disk(100, W - 40, 10) & ~disk(100, W - 40, 1.5)
    lab[disk(100, W - 40, 11)] = 0
    k += 1; lab[m] = k
    # relabel contiguous in raster order
    from .utils import renumber
    return renumber(lab)


def make_adversarial_tile(seed=0, H=256, W=256, C=7, sigma_flow=0.3, sigma_prob=1.0):
    rng = np.random.default_rng(4321 + seed)
    lab = adversarial_labels(H, W)
    n = int(lab.max())
    cell_class = np.zeros(n + 1, np.int64)
    cell_class[1:] = rng.integers(1, C, size=n)
    dP, cellprob, logits, cell_class = outputs_from_labels(rng, lab, C, sigma_flow, sigma_prob,
                                                           cell_class=cell_class)
    # class ties: make two channels exactly equal on one cell, and background-majority on another
    if n >= 2 and C > 2:
        m = lab == 1
        logits[1][m] = logits[2][m] = logits.max() + 1.0
        m = lab == 2
        logits[0][m] = logits.max() + 2.0
    return dict(labels=lab, dP=dP, cellprob=cellprob, logits=logits, cell_class=cell_class)
