"""Next row N3: device-resident hand-off from the network.

The reference copies the network output to the host (`core._from_device`, core.py:37-48, 67-68), blends the
sub-tiles there (core.py:197-231) and Cellpose uploads the result again for the dynamics.  `eval_tail` takes the
network's sub-tile outputs while they are still on the GPU and runs the whole tail of `ClassposeModel.eval`
(models.py:750-770) without leaving the device: un-flip + taper blend of flows and of class logits (with the crop
of core.py:226-229), then masks and per-cell classes.
"""
from __future__ import annotations

import numpy as np
import torch

from . import transforms as btf
from .engine import get_engine


def tile_layout(Ly0, Lx0, bsize=256, augment=False, tile_overlap=0.1):
    """Padding and sub-tile geometry run_net derives for an (Ly0, Lx0) image (core.py:129-149):
    returns (pads = (ypad1, ypad2, xpad1, xpad2), geometry dict of transforms.tile_geometry)."""
    pads = btf.get_pad_yx(Ly0, Lx0, min_size=(bsize, bsize))
    Ly, Lx = Ly0 + pads[0] + pads[1], Lx0 + pads[2] + pads[3]
    return pads, btf.tile_geometry(Ly, Lx, bsize, augment=augment, tile_overlap=tile_overlap)


def eval_tail(y_tiles, nclasses, pads, geo, augment=False, device=None, want_class_masks=False, **params):
    """y_tiles: network output for B images, [B, ntiles, nclasses + 3, ly, lx] CUDA tensor laid out as the network
    emits it (class logits first, then dY, dX, cellprob: vit_sam.py:178-197).  Returns device tensors
    (masks int32 [B,H,W], counts [B], cell_class [B,LC], class_masks | None, dP [B,2,H,W], cellprob [B,H,W])."""
    eng = get_engine(device if not (isinstance(y_tiles, torch.Tensor) and y_tiles.is_cuda) else y_tiles.device)
    y = eng._dev(y_tiles, torch.float32)
    B, nt, nch, ly, lx = y.shape
    assert nch == nclasses + 3 and nt == len(geo["y0"])
    flows = y[:, :, nclasses:].contiguous()
    logits = y[:, :, :nclasses].contiguous() if nclasses > 0 else None
    ty, tx = btf.taper_1d(ly, lx)
    x4, cover = btf.tile_cover(geo["y0"], geo["x0"], ly, lx, geo["Ly"], geo["Lx"])
    g = [eng._dev(geo[k], torch.int32) for k in ("y0", "x0", "flip")]
    tyd, txd = eng._dev(ty, torch.float64), eng._dev(tx, torch.float64)
    H, W = geo["Ly"] - pads[0] - pads[1], geo["Lx"] - pads[2] - pads[3]
    from ._abi import make_params
    with torch.cuda.device(eng.device):
        if x4 and lx % 4 == 0 and pads[2] % 4 == 0 and W % 64 == 0:
            # one library call: the flow-map blend is fused with the cellprob threshold (foreground list, scaled flow
            # field and zeroed labels come out of the blend itself), then the mask path and the vote
            masks, counts, cell_class, class_masks, dP, cellprob, _ = eng.calls.eval_tail(
                flows, logits, g[0], g[1], g[2], bool(augment), tyd, txd, geo["Ly"], geo["Lx"], tuple(int(p) for p in pads),
                make_params(**params), want_class_masks)
            return masks, counts, cell_class, class_masks, dP, cellprob
        yf = eng.calls.average_tiles(flows, g[0], g[1], g[2], bool(augment), tyd, txd, geo["Ly"], geo["Lx"], tuple(pads),
                                     x4, cover)
        yc = None
        if logits is not None:
            yc = eng.calls.average_tiles(logits, g[0], g[1], g[2], False, tyd, txd, geo["Ly"], geo["Lx"], tuple(pads),
                                         x4, cover)
    dP, cellprob = yf[:, :2].contiguous(), yf[:, 2].contiguous()
    masks, counts, cell_class, class_masks = eng.compute_masks_batch(dP, cellprob, yc, want_class_masks=want_class_masks,
                                                                     **params)
    return masks, counts, cell_class, class_masks, dP, cellprob


def prepare_tiles(imgs, bsize=256, augment=False, tile_overlap=0.1, device=None):
    """Next row N4: what happens to an image before the network (models.py:641-666 normalize_img with its defaults,
    core.py:129-178 pad + make_tiles with the parity flips), on the device.
    imgs: [B, Ly, Lx, nchan] (or [Ly, Lx, nchan]) float32 / uint8, numpy or CUDA tensor.
    Returns (tiles [B, ntiles, nchan, ly, lx] CUDA float32, pads, geometry)."""
    eng = get_engine(device if not (isinstance(imgs, torch.Tensor) and imgs.is_cuda) else imgs.device)
    x = eng._dev(imgs, torch.float32)
    if x.dim() == 3:
        x = x.unsqueeze(0)
    B, Ly0, Lx0, nchan = x.shape
    pads, geo = tile_layout(Ly0, Lx0, bsize, augment=augment, tile_overlap=tile_overlap)
    tiles, lowhigh, code = eng.prepare_tiles(x.contiguous(), pads, geo["y0"], geo["x0"], geo["flip"], geo["ly"], geo["lx"])
    return tiles, pads, geo
