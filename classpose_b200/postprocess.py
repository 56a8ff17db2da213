"""Next row N1: the per-cell work of the reference's PostProcessor
(/root/reference/src/classpose/entrypoints/predict_wsi.py:595-656) on the device.

The reference loops over every cell on the host: `ndimage.find_objects`, `cv2.findContours(cell_mask,
RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)[0]`, then a shapely polygon for validity / area / perimeter / centroid.
`cell_table` gets the same quantities for a whole batch of tiles from one kernel sequence
(`cpb_cell_contours_device`): contour points identical to cv2's, polygon measures from exact integer sums.
`cells_as_reference_dicts` formats them like the reference's `curr_cell` dictionaries (slide coordinates =
points * prediction_to_slide_scale + tile origin).
"""
from __future__ import annotations

import uuid

import numpy as np
import torch

from .engine import get_engine


def cell_table(masks, counts=None, device=None, points_cap=None):
    """masks int32 [B,H,W] (CUDA tensor or numpy) -> dict of host numpy arrays:
    npoints [B,L], offsets [B,L], points [total,2] (x, y) tile pixels, area_px, bbox (ymin,ymax,xmin,xmax inclusive),
    poly_area, perimeter, centroid (x, y), valid; L = max label + 1."""
    eng = get_engine(device if not (isinstance(masks, torch.Tensor) and masks.is_cuda) else masks.device)
    m = eng._dev(masks, torch.int32)
    if m.dim() == 2:
        m = m.unsqueeze(0)
    top = int(m.max().item()) if counts is None else int(torch.as_tensor(counts).max().item())
    lcap = top + 2
    out = eng.cell_contours(m, lcap, points_cap)
    total = int(out["total"][0].item())
    if total > out["points"].shape[0]:                       # buffer was too small: one retry with the exact size
        out = eng.cell_contours(m, lcap, total)
    h = {k: v.cpu().numpy() for k, v in out.items()}
    feat = h["feat"]
    a2 = feat[..., 5].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        cx = np.where(a2 != 0, feat[..., 6] / (3.0 * a2), np.nan)
        cy = np.where(a2 != 0, feat[..., 7] / (3.0 * a2), np.nan)
    return dict(npoints=h["npoints"], offsets=h["offsets"], points=h["points"][:total], area_px=feat[..., 0],
                bbox=feat[..., 1:5], poly_area=np.abs(a2) / 2.0, perimeter=h["perimeter"],
                centroid=np.stack([cx, cy], -1), valid=h["valid"].astype(bool))


def cells_as_reference_dicts(table, cell_class, batch_coords, prediction_to_slide_scale, labels=None, colormap=None):
    """One list of cell dictionaries per tile, keys as in the reference (predict_wsi.py:643-652).  Cells whose ring has
    fewer than 4 points or is not a valid polygon are skipped and counted, as the reference does."""
    out, n_invalid = [], 0
    s = float(prediction_to_slide_scale)
    for b, origin in enumerate(batch_coords):
        origin = np.asarray(origin, np.float64)
        cells = []
        for l in np.nonzero(table["npoints"][b] > 0)[0]:
            n, off = int(table["npoints"][b, l]), int(table["offsets"][b, l])
            if n < 4 or not table["valid"][b, l]:
                n_invalid += 1
                continue
            pts = table["points"][off:off + n].astype(np.float64) * s + origin
            coords = pts.tolist()
            coords.append(list(coords[0]))
            if cell_class is not None:
                cl = int(cell_class[b][l])
                label = labels[cl - 1] if labels is not None else str(cl)     # class 0 indexes labels[-1], as upstream
                color = colormap[cl - 1] if colormap is not None else None
                class_int = cl - 1
            else:
                label, color, class_int = "cell", [0, 168, 132], 0
            c = table["centroid"][b, l] * s + origin
            cells.append({"id": str(uuid.uuid4()), "coords": coords, "class_int": class_int,
                          "area": float(table["poly_area"][b, l]) * s * s, "label": label, "color": color,
                          "perimeter": float(table["perimeter"][b, l]) * s, "centroid": np.round(c, 2).tolist()})
        out.append(cells)
    return out, n_invalid
