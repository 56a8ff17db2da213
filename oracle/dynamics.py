"""Oracle restatement of ``cellpose.dynamics`` (cellpose==4.0.8) -- 2-D path only.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the upstream
source is not on disk; this follows its published op sequence using the same
library primitives so that library semantics carry the detail:

* ``follow_flows`` / ``steps_interp``  -> real ``torch.nn.functional.grid_sample``
* ``get_masks``                         -> integer histogram, separable max-pool
* ``remove_bad_flow_masks``/``flow_error``/``masks_to_flows`` -> ``scipy.ndimage``
  ``find_objects`` / ``mean`` and a float64 Jacobi diffusion

Reference call sites this stands in for:
  /root/reference/src/classpose/models.py:120,149-159  (resize_and_compute_masks)
SURVEY.md Appendix A.1-A.5 holds the op-by-op description this file follows.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.ndimage import find_objects, mean as nd_mean

from . import utils as outils

RPAD = 20          # get_masks_torch rpad
SEED_MIN = 10      # seeds need h > 10
GROW_MIN = 2       # region growing allowed where h > 2
GROW_ITERS = 5
FLOW_SCALE = 5.0   # network flows are 5x unit vectors
DIFFUSE_LOG = False  # cellpose < 3 took log(1 + T) before the gradient; 3.x / 4.x do not


# --------------------------------------------------------------------------
# A.3 follow_flows / steps_interp
# --------------------------------------------------------------------------
def steps_interp(dP: np.ndarray, inds, niter: int) -> torch.Tensor:
    """Euler integration of pixel positions, bilinear sampling via grid_sample.

    dP: [2,H,W] float32 (dY, dX), already masked and divided by 5.
    inds: (ys, xs) of the pixels to follow.  Returns float tensor [2, npts] (y, x).
    """
    H, W = dP.shape[1:]
    npts = len(inds[0])
    pt = torch.zeros((1, 1, npts, 2), dtype=torch.float32)
    im = torch.zeros((1, 2, H, W), dtype=torch.float32)
    # grid_sample wants (x, y) order: channel/coord 0 is X
    pt[0, 0, :, 0] = torch.from_numpy(np.asarray(inds[1])).to(torch.float32)
    pt[0, 0, :, 1] = torch.from_numpy(np.asarray(inds[0])).to(torch.float32)
    im[0, 0] = torch.from_numpy(np.ascontiguousarray(dP[1])).to(torch.float32)
    im[0, 1] = torch.from_numpy(np.ascontiguousarray(dP[0])).to(torch.float32)
    shape = np.array([W, H]).astype("float") - 1
    for k in range(2):
        im[:, k] *= 2.0 / shape[k]
        pt[..., k] /= shape[k]
    pt *= 2
    pt -= 1
    for _ in range(niter):
        dPt = torch.nn.functional.grid_sample(im, pt, align_corners=False)
        for k in range(2):
            pt[..., k] = torch.clamp(pt[..., k] + dPt[:, k], -1.0, 1.0)
    pt += 1
    pt *= 0.5
    for k in range(2):
        pt[..., k] *= shape[k]
    out = pt[..., [1, 0]].squeeze()
    if out.ndim == 1:
        out = out.unsqueeze(0)
    return out.T


def follow_flows(dP: np.ndarray, inds, niter: int = 200) -> torch.Tensor:
    return steps_interp(dP, inds, niter)


# --------------------------------------------------------------------------
# A.4 get_masks
# --------------------------------------------------------------------------
def max_pool1d(h: np.ndarray, kernel_size: int, axis: int) -> np.ndarray:
    """stride-1 max pool with edge-truncated windows along one axis."""
    out = h.copy()
    nd = h.shape[axis]
    k0 = kernel_size // 2
    for d in range(-k0, k0 + 1):
        dst = [slice(None)] * h.ndim
        src = [slice(None)] * h.ndim
        dst[axis] = slice(max(-d, 0), min(nd - d, nd))
        src[axis] = slice(max(d, 0), min(nd + d, nd))
        np.maximum(out[tuple(dst)], h[tuple(src)], out=out[tuple(dst)])
    return out


def max_pool_nd(h: np.ndarray, kernel_size: int = 5) -> np.ndarray:
    """2-D separable max pool over the last two axes (leading axis is batch)."""
    return max_pool1d(max_pool1d(h, kernel_size, 1), kernel_size, 2)


def get_masks(p_final: np.ndarray, inds, shape0, rpad: int = RPAD,
              max_size_fraction: float = 0.4) -> np.ndarray:
    """Histogram of end points -> seeds -> 11x11 constrained growth -> labels.

    p_final: int [2, npts] truncated end positions (y, x).  Returns uint16/uint32 [H,W].

    Tie rule: upstream sorts seeds with an *unstable* argsort, so equal-count
    seeds paint in an arbitrary order there.  The oracle fixes the rule the
    CUDA path reproduces: stable ascending sort, i.e. among equal counts the
    seed later in raster order paints last (wins overlaps).
    """
    shape0 = tuple(int(s) for s in shape0)
    pt = np.asarray(p_final).astype(np.int64) + rpad
    pt = np.maximum(pt, 0)
    for i in range(2):
        pt[i] = np.minimum(pt[i], shape0[i] + rpad - 1)
    shape = tuple(np.array(shape0) + 2 * rpad)

    h1 = np.zeros(shape, np.int32)
    np.add.at(h1, (pt[0], pt[1]), 1)

    hmax1 = max_pool_nd(h1[None], kernel_size=5)[0]
    seeds1 = np.stack(np.nonzero((h1 - hmax1 > -1e-6) & (h1 > SEED_MIN)), axis=1)
    if len(seeds1) == 0:
        return np.zeros(shape0, dtype="uint16")
    npts = h1[tuple(seeds1.T)]
    isort1 = np.argsort(npts, kind="stable")
    seeds1 = seeds1[isort1]

    n_seeds = len(seeds1)
    h_slc = np.zeros((n_seeds, 11, 11), np.float32)
    for k in range(n_seeds):
        sy, sx = seeds1[k]
        h_slc[k] = h1[sy - 5:sy + 6, sx - 5:sx + 6]
    seed_masks = np.zeros((n_seeds, 11, 11), np.float32)
    seed_masks[:, 5, 5] = 1
    for _ in range(GROW_ITERS):
        seed_masks = max_pool_nd(seed_masks, kernel_size=3)
        seed_masks *= h_slc > GROW_MIN

    dtype = np.int32 if n_seeds < 2 ** 16 else np.int64
    M1 = np.zeros(shape, dtype)
    for k in range(n_seeds):
        yy, xx = np.nonzero(seed_masks[k])
        M1[yy + seeds1[k][0] - 5, xx + seeds1[k][1] - 5] = 1 + k

    lab = M1[pt[0], pt[1]]
    M0 = np.zeros(shape0, dtype="uint16" if n_seeds < 2 ** 16 else "uint32")
    M0[inds] = lab

    uniq, counts = np.unique(M0, return_counts=True)
    big = np.prod(shape0) * max_size_fraction
    bigc = uniq[counts > big]
    if len(bigc) > 0 and (len(bigc) > 1 or bigc[0] != 0):
        M0 = outils.mask_labels(M0, bigc)
    M0 = outils.renumber(M0)
    return M0.reshape(shape0)


# --------------------------------------------------------------------------
# A.5 masks_to_flows / flow_error / remove_bad_flow_masks
# --------------------------------------------------------------------------
_NEIGH_Y = np.array([0, -1, 1, 0, 0, -1, -1, 1, 1])
_NEIGH_X = np.array([0, 0, 0, -1, 1, -1, 1, -1, 1])


def get_centers(masks: np.ndarray, slices):
    """Per label: the label pixel nearest to the label's mean position (first
    arg-min in raster order of the bbox crop) and ext = bbox_h + bbox_w + 2."""
    centers = np.zeros((len(slices), 2), "int32")
    ext = np.zeros((len(slices),), "int32")
    for i, si in enumerate(slices):
        if si is None:
            continue
        sr, sc = si
        yi, xi = np.nonzero(masks[sr, sc] == (i + 1))
        yi = yi.astype(np.int32) + 1
        xi = xi.astype(np.int32) + 1
        ymed = yi.mean()
        xmed = xi.mean()
        imin = ((xi - xmed) ** 2 + (yi - ymed) ** 2).argmin()
        centers[i, 0] = yi[imin] + sr.start - 1
        centers[i, 1] = xi[imin] + sc.start - 1
        ext[i] = (sr.stop - sr.start + 1) + (sc.stop - sc.start + 1)
    return centers, ext


def extend_centers(neighbors, centers, isneighbor, shape, n_iter: int) -> np.ndarray:
    """float64 Jacobi heat diffusion from the centres; returns raw (dy, dx) [2, npix]."""
    if np.prod(shape) > 4e7:
        T = np.zeros(shape, np.float32)
    else:
        T = np.zeros(shape, np.float64)
    cy, cx = centers[:, 0], centers[:, 1]
    ny, nx = neighbors
    for _ in range(int(n_iter)):
        T[cy, cx] += 1
        Tneigh = T[ny, nx]
        Tneigh *= isneighbor
        T[ny[0], nx[0]] = Tneigh.mean(axis=0)
    if DIFFUSE_LOG:
        T = np.log(1.0 + T)
    dy = T[ny[2], nx[2]] - T[ny[1], nx[1]]
    dx = T[ny[4], nx[4]] - T[ny[3], nx[3]]
    return np.stack((dy, dx), axis=0)


def masks_to_flows(masks: np.ndarray, niter: int | None = None) -> np.ndarray:
    """Unit flow field [2,H,W] float64 derived from a label image by diffusion."""
    Ly0, Lx0 = masks.shape
    if masks.max() <= 0:
        return np.zeros((2, Ly0, Lx0))
    mp = np.pad(masks.astype(np.int64), 1)
    y, x = np.nonzero(mp)
    neighbors = np.stack((y[None] + _NEIGH_Y[:, None], x[None] + _NEIGH_X[:, None]), axis=0)
    m0 = mp[neighbors[0, 0], neighbors[1, 0]]
    isneighbor = mp[neighbors[0], neighbors[1]] == m0[None]
    slices = find_objects(masks)
    centers, ext = get_centers(masks, slices)
    centers = centers.astype(np.int64) + 1
    n_iter = 2 * int(ext.max()) if niter is None else niter
    mu = extend_centers(neighbors, centers, isneighbor, mp.shape, n_iter).astype("float64")
    mu /= (1e-60 + (mu ** 2).sum(axis=0) ** 0.5)
    mu0 = np.zeros((2, Ly0, Lx0))
    mu0[:, y - 1, x - 1] = mu
    return mu0


def flow_error(maski: np.ndarray, dP_net: np.ndarray):
    dP_masks = masks_to_flows(maski)
    nlab = int(maski.max())
    flow_errors = np.zeros(nlab)
    for i in range(dP_masks.shape[0]):
        flow_errors += nd_mean((dP_masks[i] - dP_net[i] / FLOW_SCALE) ** 2, maski,
                               index=np.arange(1, nlab + 1))
    return flow_errors, dP_masks


def remove_bad_flow_masks(masks: np.ndarray, flows: np.ndarray, threshold: float = 0.4):
    merrors, _ = flow_error(masks, flows)
    badi = 1 + (merrors > threshold).nonzero()[0]
    masks[np.isin(masks, badi)] = 0
    return masks


# --------------------------------------------------------------------------
# A.2 compute_masks / A.1 resize_and_compute_masks
# --------------------------------------------------------------------------
def compute_masks(dP, cellprob, niter=200, cellprob_threshold=0.0, flow_threshold=0.4,
                  do_3D=False, min_size=-1, max_size_fraction=0.4, device=None,
                  return_stages: dict | None = None):
    """2-D restatement of cellpose.dynamics.compute_masks.  `return_stages`, when a
    dict, receives intermediate results for stage-wise parity tests."""
    assert not do_3D, "oracle covers the 2-D path only"
    dP = np.asarray(dP, np.float32)
    cellprob = np.asarray(cellprob, np.float32)
    fg = cellprob > cellprob_threshold
    if fg.sum() == 0:
        return np.zeros(cellprob.shape, "uint16")
    inds = np.nonzero(fg)
    p = follow_flows(dP * fg / FLOW_SCALE, inds, niter)
    p_final = p.int().numpy()
    mask = get_masks(p_final, inds, dP.shape[1:], max_size_fraction=max_size_fraction)
    if return_stages is not None:
        return_stages["inds"] = inds
        return_stages["p_float"] = p.numpy().copy()
        return_stages["p_final"] = p_final.copy()
        return_stages["masks_get"] = mask.copy()
    if mask.max() > 0 and flow_threshold is not None and flow_threshold > 0:
        mask = remove_bad_flow_masks(mask, dP, threshold=flow_threshold)
    if return_stages is not None:
        return_stages["masks_qc"] = mask.copy()
    if mask.max() < 2 ** 16 and mask.dtype != np.uint16:
        mask = mask.astype("uint16")
    if min_size > 0:
        mask = outils.fill_holes_and_remove_small_masks(mask, min_size=min_size)
    return mask


def resize_and_compute_masks(dP, cellprob, niter=200, cellprob_threshold=0.0,
                             flow_threshold=0.4, do_3D=False, min_size=15,
                             max_size_fraction=0.4, resize=None, device=None,
                             return_stages: dict | None = None):
    """compute_masks (min_size not forwarded) then hole fill + size filter.
    `resize` is accepted and ignored, as in cellpose 4."""
    mask = compute_masks(dP, cellprob, niter=niter, cellprob_threshold=cellprob_threshold,
                         flow_threshold=flow_threshold, do_3D=do_3D,
                         max_size_fraction=max_size_fraction, device=device,
                         return_stages=return_stages)
    mask = outils.fill_holes_and_remove_small_masks(mask, min_size=min_size)
    return mask
