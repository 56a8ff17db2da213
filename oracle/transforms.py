"""Oracle restatement of the ``cellpose.transforms`` tile helpers used by
``classpose.core.run_net`` (cellpose==4.0.8; not on disk -> PARITY UNPINNED).

TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md Appendix A.7.
Reference call sites: /root/reference/src/classpose/core.py:130 (get_pad_yx),
:173 (make_tiles), :209 (unaugment_tiles), :215,218 (average_tiles).
"""
from __future__ import annotations

import numpy as np


def get_pad_yx(Ly, Lx, div=16, extra=1, min_size=None):
    if min_size is None or Ly >= min_size[-2]:
        Lpad = int(div * np.ceil(Ly / div) - Ly)
    else:
        Lpad = min_size[-2] - Ly
    ypad1 = extra * div // 2 + Lpad // 2
    ypad2 = extra * div // 2 + Lpad - Lpad // 2
    if min_size is None or Lx >= min_size[-1]:
        Lpad = int(div * np.ceil(Lx / div) - Lx)
    else:
        Lpad = min_size[-1] - Lx
    xpad1 = extra * div // 2 + Lpad // 2
    xpad2 = extra * div // 2 + Lpad - Lpad // 2
    return ypad1, ypad2, xpad1, xpad2


def _flip_code(j, i):
    """0 none, 1 flip Y, 2 flip X, 3 both -- the parity pattern of augmented tiles."""
    if j % 2 == 0 and i % 2 == 1:
        return 1
    if j % 2 == 1 and i % 2 == 0:
        return 2
    if j % 2 == 1 and i % 2 == 1:
        return 3
    return 0


def _apply_flip(a, code):
    if code == 1:
        return a[..., ::-1, :]
    if code == 2:
        return a[..., :, ::-1]
    if code == 3:
        return a[..., ::-1, ::-1]
    return a


def make_tiles(imgi, bsize=224, augment=False, tile_overlap=0.1):
    nchan, Ly, Lx = imgi.shape
    if augment:
        bsize = np.int32(bsize)
        if Ly < bsize:
            imgi = np.concatenate((imgi, np.zeros((nchan, bsize - Ly, Lx))), axis=1)
            Ly = bsize
        if Lx < bsize:
            imgi = np.concatenate((imgi, np.zeros((nchan, Ly, bsize - Lx))), axis=2)
        Ly, Lx = imgi.shape[-2:]
        ny = max(2, int(np.ceil(2. * Ly / bsize)))
        nx = max(2, int(np.ceil(2. * Lx / bsize)))
        bsizeY = bsizeX = bsize
    else:
        tile_overlap = min(0.5, max(0.05, tile_overlap))
        bsizeY, bsizeX = np.int32(min(bsize, Ly)), np.int32(min(bsize, Lx))
        ny = 1 if Ly <= bsize else int(np.ceil((1. + 2 * tile_overlap) * Ly / bsize))
        nx = 1 if Lx <= bsize else int(np.ceil((1. + 2 * tile_overlap) * Lx / bsize))
    ystart = np.linspace(0, Ly - bsizeY, ny).astype(int)
    xstart = np.linspace(0, Lx - bsizeX, nx).astype(int)
    ysub, xsub = [], []
    IMG = np.zeros((len(ystart), len(xstart), nchan, bsizeY, bsizeX), np.float32)
    for j in range(len(ystart)):
        for i in range(len(xstart)):
            ysub.append([ystart[j], ystart[j] + bsizeY])
            xsub.append([xstart[i], xstart[i] + bsizeX])
            t = imgi[:, ysub[-1][0]:ysub[-1][1], xsub[-1][0]:xsub[-1][1]]
            IMG[j, i] = _apply_flip(t, _flip_code(j, i)) if augment else t
    return IMG, ysub, xsub, Ly, Lx


def unaugment_tiles(y):
    """Undo the flips of augmented tiles; flow channel 0 (dY) changes sign on Y flips,
    channel 1 (dX) on X flips.  y: [ny, nx, 3, ly, lx]; modified in place and returned."""
    for j in range(y.shape[0]):
        for i in range(y.shape[1]):
            code = _flip_code(j, i)
            if code:
                y[j, i] = _apply_flip(y[j, i], code)
                if code & 1:
                    y[j, i, 0] *= -1
                if code & 2:
                    y[j, i, 1] *= -1
    return y


def _taper_mask(ly=224, lx=224, sig=7.5):
    bsize = max(224, max(ly, lx))
    xm = np.arange(bsize)
    xm = np.abs(xm - xm.mean())
    mask = 1 / (1 + np.exp((xm - (bsize / 2 - 20)) / sig))
    mask = mask * mask[:, np.newaxis]
    mask = mask[bsize // 2 - ly // 2:bsize // 2 + ly // 2 + ly % 2,
                bsize // 2 - lx // 2:bsize // 2 + lx // 2 + lx % 2]
    return mask


def average_tiles(y, ysub, xsub, Ly, Lx):
    """Taper-weighted average of overlapping tiles: y [ntiles, nch, ly, lx] -> [nch, Ly, Lx] f32."""
    Navg = np.zeros((Ly, Lx))
    yf = np.zeros((y.shape[1], Ly, Lx), np.float32)
    mask = _taper_mask(ly=y.shape[-2], lx=y.shape[-1])
    for j in range(len(ysub)):
        yf[:, ysub[j][0]:ysub[j][1], xsub[j][0]:xsub[j][1]] += y[j] * mask
        Navg[ysub[j][0]:ysub[j][1], xsub[j][0]:xsub[j][1]] += mask
    yf /= Navg
    return yf


# ---- tile preparation in front of the network --------------------------------------
def normalize99(Y, lower=1, upper=99):
    """cellpose.transforms.normalize99 (no down-sampling branch: a tile is far below 224**3 values)."""
    X = Y.astype("float32").copy()
    x01 = np.percentile(X, lower)
    x99 = np.percentile(X, upper)
    if x99 - x01 > 1e-3:
        X -= x01
        X /= (x99 - x01)
    else:
        X[:] = 0
    return X


def normalize_img(img, percentile=(1.0, 99.0)):
    """cellpose.transforms.normalize_img with its defaults as Classpose calls it (models.py:641-666):
    per channel (last axis) normalize99 over the whole image; constant channels are left untouched."""
    img_norm = img.astype(np.float32).copy()
    for c in range(img_norm.shape[-1]):
        if np.ptp(img_norm[..., c]) > 0.0:
            img_norm[..., c] = normalize99(img_norm[..., c], lower=percentile[0], upper=percentile[1])
    return img_norm


def prepare_tiles(img, bsize=256, augment=False, tile_overlap=0.1):
    """One [Ly, Lx, nchan] image -> the network input run_net builds (core.py:129-178):
    normalise, pad to the /16 grid, cut (and flip) the sub-tiles.  Returns (IMG [ntiles, nchan, ly, lx], ysub, xsub, pads)."""
    x = normalize_img(img)
    pads = get_pad_yx(x.shape[0], x.shape[1], min_size=(bsize, bsize))
    imgb = np.pad(x.transpose(2, 0, 1), np.array([[0, 0], [pads[0], pads[1]], [pads[2], pads[3]]]), mode="constant")
    IMG, ysub, xsub, Ly, Lx = make_tiles(imgb, bsize=bsize, augment=augment, tile_overlap=tile_overlap)
    return IMG.reshape(-1, imgb.shape[0], IMG.shape[-2], IMG.shape[-1]), ysub, xsub, pads
