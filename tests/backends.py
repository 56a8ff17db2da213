"""numpy-in / numpy-out adapters over the C-ABI call layer, so the same parity cases run against
the simulated kernels on the CPU box and against the real library on the B200."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "sim"))


def _np(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x
    return x.detach().cpu().numpy()


class _Base:
    def _calls(self):
        raise NotImplementedError

    def follow_flows(self, dP, cp, niter=200, thr=0.0, want_float=False):
        a, b = self._calls().follow_flows(self._in(dP), self._in(cp), niter, thr, want_float)
        return _np(a), _np(b)

    def get_masks(self, pf, msf=0.4):
        a, b = self._calls().get_masks(self._in(pf), msf)
        return _np(a), _np(b)

    def masks_to_flows(self, masks, lcap):
        return _np(self._calls().masks_to_flows(self._in(masks), lcap))

    def remove_bad_flow_masks(self, masks, dP, lcap, thr=0.4, want_err=False):
        a, b = self._calls().remove_bad_flow_masks(self._in(masks), self._in(dP), lcap, thr, want_err)
        return _np(a), _np(b)

    def fill_holes_and_remove_small_masks(self, masks, lcap, min_size=15):
        a, b = self._calls().fill_holes_and_remove_small_masks(self._in(masks), lcap, min_size)
        return _np(a), _np(b)

    def class_vote(self, masks, logits, lcap, want_class_masks=True):
        a, b = self._calls().class_vote(self._in(masks), self._in(logits), lcap, want_class_masks)
        return _np(a), _np(b)

    def remove_border_instances(self, masks, lcap, nch=1):
        return _np(self._calls().remove_border_instances(self._in(masks), lcap, nch))

    def average_tiles(self, y, y0, x0, flip, negate, ty, tx, Ly, Lx, crop=(0, 0, 0, 0), vector=True):
        from classpose_b200.transforms import tile_cover
        x4, cover = tile_cover(y0, x0, y.shape[-2], y.shape[-1], Ly, Lx) if vector else (False, 0)
        return _np(self._calls().average_tiles(self._in(y), self._in(np.asarray(y0, np.int32)),
                                               self._in(np.asarray(x0, np.int32)), self._in(np.asarray(flip, np.int32)),
                                               negate, self._in(np.asarray(ty, np.float64)),
                                               self._in(np.asarray(tx, np.float64)), Ly, Lx, crop, x4, cover))

    def compute_masks(self, dP, cp, logits=None, want_class_masks=False, **kw):
        from classpose_b200._abi import make_params
        prm = make_params(**kw)
        out = self._calls().compute_masks(self._in(dP), self._in(cp), self._in(logits), prm, want_class_masks)
        return tuple(_np(o) for o in out)

    def eval_tail(self, y_flows, y_logits, y0, x0, flip, augment, ty, tx, Ly, Lx, crop, want_class_masks=False, **kw):
        from classpose_b200._abi import make_params
        out = self._calls().eval_tail(self._in(y_flows), self._in(y_logits), self._in(np.asarray(y0, np.int32)),
                                      self._in(np.asarray(x0, np.int32)), self._in(np.asarray(flip, np.int32)), augment,
                                      self._in(np.asarray(ty, np.float64)), self._in(np.asarray(tx, np.float64)), Ly, Lx, crop,
                                      make_params(**kw), want_class_masks)
        return tuple(_np(o) for o in out)

    def cell_contours(self, masks, lcap, points_cap=None):
        out = self._calls().cell_contours(self._in(masks), lcap, points_cap)
        return {k: _np(v) for k, v in out.items()}

    def dedup_cells(self, cx, cy, size, max_dist=7.5, want_group=False):
        a, b = self._calls().dedup_cells(self._in(np.asarray(cx, np.float64)), self._in(np.asarray(cy, np.float64)),
                                         self._in(np.asarray(size, np.float64)), max_dist, want_group)
        return _np(a), _np(b)

    def prepare_tiles(self, img, pads, y0, x0, flip, ly, lx, lower=1.0, upper=99.0):
        out = self._calls().prepare_tiles(self._in(img), pads, self._in(np.asarray(y0, np.int32)),
                                          self._in(np.asarray(x0, np.int32)), self._in(np.asarray(flip, np.int32)), ly, lx,
                                          lower, upper)
        return tuple(_np(o) for o in out)

    def set_follow_merge(self, mode):
        self._calls().lib.cpb_debug_set_follow_merge(int(mode))

    def set_switch(self, which, value):
        self._calls().lib.cpb_debug_set_switch(int(which), int(value))

    def label_offsets(self, counts, base=0):
        a, b = self._calls().label_offsets(self._in(counts), base)
        return _np(a), _np(b)


class SimBackend(_Base):
    name = "sim"

    def __init__(self):
        import simlib
        self.c = simlib.calls()

    def _calls(self):
        return self.c

    def _in(self, x):
        return None if x is None else np.ascontiguousarray(x)


class GpuBackend(_Base):
    name = "cuda"

    def __init__(self):
        import torch
        from classpose_b200.engine import get_engine
        self.torch = torch
        self.eng = get_engine()

    def _calls(self):
        return self.eng.calls

    def _in(self, x):
        if x is None:
            return None
        return self.torch.from_numpy(np.ascontiguousarray(x)).to(self.eng.device)
