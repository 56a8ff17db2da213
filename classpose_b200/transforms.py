"""Drop-in for the `cellpose.transforms` / `classpose.transforms` tile helpers on the hot path
(hook D of SURVEY.md 8b), plus the host-side tile geometry run_net derives before blending.

  average_tiles          <- cellpose.transforms.average_tiles, called at
                            /root/reference/src/classpose/core.py:215, 218-220
  unaugment_tiles        <- cellpose.transforms.unaugment_tiles (core.py:209)
  unaugment_class_tiles  <- /root/reference/src/classpose/transforms/transforms.py:4-21
  get_pad_yx, tile_geometry, taper_1d  <- core.py:130-149 and make_tiles / _taper_mask upstream

The blend itself (taper-weighted accumulate, un-flip, flow sign change, crop) is one CUDA
kernel; the geometry is tiny integer host logic.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import get_engine


def get_pad_yx(Ly, Lx, div=16, extra=1, min_size=None):
    def one(L, m):
        Lpad = int(div * np.ceil(L / div) - L) if (m is None or L >= m) else m - L
        return extra * div // 2 + Lpad // 2, extra * div // 2 + Lpad - Lpad // 2
    y1, y2 = one(Ly, None if min_size is None else min_size[-2])
    x1, x2 = one(Lx, None if min_size is None else min_size[-1])
    return y1, y2, x1, x2


def flip_code(j, i):
    """bit0: tile was flipped in Y, bit1: flipped in X (the parity pattern of augmented tiles)."""
    if j % 2 == 0 and i % 2 == 1:
        return 1
    if j % 2 == 1 and i % 2 == 0:
        return 2
    if j % 2 == 1 and i % 2 == 1:
        return 3
    return 0


def tile_geometry(Ly, Lx, bsize=224, augment=False, tile_overlap=0.1):
    """Window origins, size and flip codes of the sub-tiles make_tiles would cut from an (Ly, Lx) image.
    Returns dict(ystart, xstart, ly, lx, Ly, Lx, y0[ntiles], x0[ntiles], flip[ntiles], ny, nx)."""
    if augment:
        Ly, Lx = max(Ly, bsize), max(Lx, bsize)
        ny = max(2, int(np.ceil(2.0 * Ly / bsize)))
        nx = max(2, int(np.ceil(2.0 * Lx / bsize)))
        ly = lx = int(bsize)
    else:
        tile_overlap = min(0.5, max(0.05, tile_overlap))
        ly, lx = int(min(bsize, Ly)), int(min(bsize, Lx))
        ny = 1 if Ly <= bsize else int(np.ceil((1.0 + 2 * tile_overlap) * Ly / bsize))
        nx = 1 if Lx <= bsize else int(np.ceil((1.0 + 2 * tile_overlap) * Lx / bsize))
    ystart = np.linspace(0, Ly - ly, ny).astype(int)
    xstart = np.linspace(0, Lx - lx, nx).astype(int)
    y0 = np.repeat(ystart, nx).astype(np.int32)
    x0 = np.tile(xstart, ny).astype(np.int32)
    flip = np.array([flip_code(j, i) if augment else 0 for j in range(ny) for i in range(nx)], np.int32)
    return dict(ystart=ystart, xstart=xstart, ly=ly, lx=lx, Ly=int(Ly), Lx=int(Lx), y0=y0, x0=x0, flip=flip,
                ny=ny, nx=nx)


def taper_1d(ly, lx, sig=7.5):
    """The two 1-D factors of cellpose's _taper_mask (its 2-D mask is their outer product), float64."""
    bsize = max(224, max(ly, lx))
    xm = np.arange(bsize)
    xm = np.abs(xm - xm.mean())
    m = 1 / (1 + np.exp((xm - (bsize / 2 - 20)) / sig))
    ty = m[bsize // 2 - ly // 2: bsize // 2 + ly // 2 + ly % 2]
    tx = m[bsize // 2 - lx // 2: bsize // 2 + lx // 2 + lx % 2]
    return np.ascontiguousarray(ty), np.ascontiguousarray(tx)


def tile_cover(y0, x0, ly, lx, Ly, Lx):
    """Host-side facts about a window layout that let the blend use its 128-bit kernel:
    (every x0 is a multiple of 4, maximum number of windows covering one pixel)."""
    y0, x0 = np.asarray(y0, np.int64), np.asarray(x0, np.int64)
    cover = np.zeros((int(Ly), int(Lx)), np.int32)
    for a, b in zip(y0, x0):
        cover[a:a + ly, b:b + lx] += 1
    return bool((x0 % 4 == 0).all()), int(cover.max())


def blend_tiles(y, y0, x0, Ly, Lx, flip=None, negate_flow=False, crop=(0, 0, 0, 0), device=None):
    """Batched blend: y [B,ntiles,nch,ly,lx] (numpy or CUDA tensor) -> [B,nch,Ly-crop,Lx-crop].
    `flip`/`negate_flow` fuse unaugment_tiles (flows) or unaugment_class_tiles (logits)."""
    eng = get_engine(device if not (isinstance(y, torch.Tensor) and y.is_cuda) else y.device)
    B, ntiles, nch, ly, lx = y.shape
    ty, tx = taper_1d(ly, lx)
    flip = np.zeros(ntiles, np.int32) if flip is None else np.asarray(flip, np.int32)
    x4, cover = tile_cover(y0, x0, ly, lx, Ly, Lx)
    out = eng.average_tiles(y, np.asarray(y0, np.int32), np.asarray(x0, np.int32), flip, negate_flow, ty, tx,
                            int(Ly), int(Lx), tuple(int(c) for c in crop), x0_multiple_of_4=x4, max_cover=cover)
    return out if (isinstance(y, torch.Tensor) and y.is_cuda) else out.cpu().numpy()


def average_tiles(y, ysub, xsub, Ly, Lx, device=None):
    """cellpose.transforms.average_tiles: y [ntiles,nch,ly,lx], windows ysub/xsub -> [nch,Ly,Lx] float32."""
    y = np.ascontiguousarray(np.asarray(y, np.float32))
    y0 = np.array([s[0] for s in ysub], np.int32)
    x0 = np.array([s[0] for s in xsub], np.int32)
    return blend_tiles(y[None], y0, x0, Ly, Lx, device=device)[0]


def _unaugment(y, negate_flow, device):
    """Un-flip tiles [ny,nx,nch,ly,lx]: a blend of a single tile onto its own window with unit weights
    is exactly the un-flip (v*1.0 / 1.0), so the blend kernel does it, one launch per flip code."""
    ny, nx, nch, ly, lx = y.shape
    flip = np.array([flip_code(j, i) for j in range(ny) for i in range(nx)], np.int32)
    yy = np.ascontiguousarray(np.asarray(y, np.float32)).reshape(ny * nx, 1, nch, ly, lx)
    eng = get_engine(device)
    ones_y, ones_x, zero = np.ones(ly), np.ones(lx), np.zeros(1, np.int32)
    res = np.empty((ny * nx, nch, ly, lx), np.float32)
    for code in range(4):
        sel = np.nonzero(flip == code)[0]
        if len(sel):
            part = eng.average_tiles(yy[sel], zero, zero, np.array([code], np.int32), negate_flow, ones_y, ones_x,
                                     ly, lx)
            res[sel] = part.cpu().numpy()
    return res.reshape(ny, nx, nch, ly, lx)


def unaugment_tiles(y, device=None):
    """cellpose.transforms.unaugment_tiles: undo flips; dY changes sign on Y flips, dX on X flips."""
    return _unaugment(y, True, device)


def unaugment_class_tiles(y, device=None):
    """classpose.transforms.unaugment_class_tiles: undo flips only."""
    if isinstance(y, torch.Tensor):
        return torch.from_numpy(_unaugment(y.cpu().numpy(), False, device))
    return _unaugment(y, False, device)
