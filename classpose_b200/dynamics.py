"""Drop-in for the `cellpose.dynamics` functions Classpose calls (hook B of SURVEY.md 8b):
same names, argument meaning and "no foreground -> zeros" behaviour, computed by the sm_100a
kernels.  Call sites replaced: /root/reference/src/classpose/models.py:120, 149-159.

numpy in -> numpy out (uint16, or uint32 from 65536 labels, as Cellpose returns);
CUDA torch tensors in -> CUDA int32 tensors out (no host round trip).
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from ._abi import ClassposeB200Error
from .engine import get_engine


def _is_dev(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def _label_dtype(m: np.ndarray):
    return np.uint16 if (m.size == 0 or m.max() < 2 ** 16) else np.uint32


def _as_batch(dP, cellprob):
    if isinstance(dP, torch.Tensor):
        dP_b = dP.unsqueeze(0) if dP.dim() == 3 else dP
        cp_b = cellprob.unsqueeze(0) if cellprob.dim() == 2 else cellprob
    else:
        dP = np.asarray(dP, np.float32)
        cellprob = np.asarray(cellprob, np.float32)
        dP_b = dP[None] if dP.ndim == 3 else dP
        cp_b = cellprob[None] if cellprob.ndim == 2 else cellprob
    if dP_b.shape[1] != 2 or tuple(dP_b.shape[2:]) != tuple(cp_b.shape[1:]) or dP_b.shape[0] != cp_b.shape[0]:
        raise ValueError(f"dP {tuple(dP.shape)} and cellprob {tuple(cellprob.shape)} do not describe the same 2-D tile")
    return dP_b, cp_b


def _run(dP, cellprob, niter, cellprob_threshold, flow_threshold, min_size, max_size_fraction, fill_holes, device):
    eng = get_engine(device if _is_dev(dP) is False else dP.device)
    dP_b, cp_b = _as_batch(dP, cellprob)
    if not isinstance(dP_b, torch.Tensor) and dP_b.shape[0] == 1 and (dP_b.shape[2] * dP_b.shape[3]) % 4 == 0:
        # one tile of host data (the reference's WSI loop): static buffers + one CUDA-graph launch (fastpath.py)
        from . import fastpath
        plan = fastpath.tile_plan(eng, dP_b.shape[2], dP_b.shape[3], niter, cellprob_threshold, flow_threshold, min_size,
                                  max_size_fraction, fill_holes)
        m32, n = plan.run(dP_b[0], cp_b[0])
        m = m32.astype(np.uint16 if n < 2 ** 16 else np.uint32)
        plan.remember(m)                 # hook C finds the labels still on the device when it is handed this array
        return m if np.ndim(dP) == 3 else m[None]
    masks, counts, _, _ = eng.compute_masks_batch(dP_b, cp_b, None, niter=niter, cellprob_threshold=cellprob_threshold,
                                                  flow_threshold=flow_threshold, min_size=min_size,
                                                  max_size_fraction=max_size_fraction, fill_holes=fill_holes)
    single = (dP.dim() if isinstance(dP, torch.Tensor) else np.ndim(dP)) == 3
    if _is_dev(dP):
        return masks[0] if single else masks      # asynchronous: a failed tile shows as counts[b] = -1 (engine API)
    if bool((counts < 0).any().item()):
        raise ClassposeB200Error("a tile exhausted the hole-fill bitmap pool (counts = -1): masks are incomplete")
    m = masks.cpu().numpy()
    m = m.astype(_label_dtype(m))
    return m[0] if single else m


def compute_masks(dP, cellprob, p=None, niter=200, cellprob_threshold=0.0, flow_threshold=0.4, do_3D=False,
                  min_size=-1, max_size_fraction=0.4, device=None):
    """cellpose.dynamics.compute_masks, 2-D: follow flows -> masks -> flow-error check; hole fill and
    size filter only when min_size > 0 (its default here is -1, as upstream)."""
    if do_3D:
        raise NotImplementedError("classpose_b200 covers the 2-D post-network path only")
    if p is not None:
        raise NotImplementedError("pre-computed pixel positions `p` are not supported")
    return _run(dP, cellprob, niter, cellprob_threshold, flow_threshold, min_size, max_size_fraction,
                fill_holes=min_size > 0, device=device)


def resize_and_compute_masks(dP, cellprob, niter=200, cellprob_threshold=0.0, flow_threshold=0.4, do_3D=False,
                             min_size=15, max_size_fraction=0.4, resize=None, device=None):
    """cellpose.dynamics.resize_and_compute_masks: compute_masks, then fill holes and drop small masks.
    `resize` is accepted and ignored with a warning, as in cellpose 4."""
    if do_3D:
        raise NotImplementedError("classpose_b200 covers the 2-D post-network path only")
    if resize is not None:
        warnings.warn("resize is deprecated in cellpose 4 and ignored", stacklevel=2)
    return _run(dP, cellprob, niter, cellprob_threshold, flow_threshold, min_size, max_size_fraction,
                fill_holes=True, device=device)


def follow_flows(dP, inds=None, niter=200, device=None, cellprob=None, cellprob_threshold=0.0):
    """Euler-integrate pixels through `dP` (already masked and scaled by 1/5 upstream; here pass the raw
    network dP plus `cellprob`, the masking and scaling happen on the device).  Returns float [2, npts]
    (y, x) for the pixels `inds` (default: all foreground pixels, raster order) -- numpy in, numpy out."""
    if cellprob is None:
        raise ClassposeB200Error("follow_flows needs `cellprob` (masking and /5 are fused into the kernel)")
    eng = get_engine(device)
    dP_b, cp_b = _as_batch(dP, cellprob)
    _, pfl = eng.follow_flows(dP_b, cp_b, niter, cellprob_threshold, want_float=True)
    pfl = pfl[0].cpu().numpy()
    if inds is None:
        inds = np.nonzero(np.asarray(cellprob) > cellprob_threshold)
    return pfl[:, inds[0], inds[1]]


def masks_to_flows(masks, device=None, niter=None):
    """cellpose.dynamics.masks_to_flows: float64 unit flows [2,H,W] from a label image."""
    if niter is not None:
        raise NotImplementedError("fixed niter is not supported; the reference path never passes it")
    eng = get_engine(device)
    m = np.ascontiguousarray(np.asarray(masks).astype(np.int32))[None]
    if m.max() <= 0:
        return np.zeros((2,) + m.shape[1:])
    return eng.masks_to_flows(m, int(m.max()) + 2)[0].cpu().numpy()


def remove_bad_flow_masks(masks, flows, threshold=0.4, device=None):
    """cellpose.dynamics.remove_bad_flow_masks: zero labels whose flow error exceeds `threshold`."""
    eng = get_engine(device)
    m = np.asarray(masks)
    if m.max() <= 0:
        return masks
    out, _ = eng.remove_bad_flow_masks(np.ascontiguousarray(m.astype(np.int32))[None],
                                       np.ascontiguousarray(np.asarray(flows, np.float32))[None],
                                       int(m.max()) + 2, threshold)
    return out[0].cpu().numpy().astype(m.dtype)
