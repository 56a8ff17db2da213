"""Thin, memory-agnostic call layer over the C ABI.

`Calls` owns no memory policy: it is handed a `mem` object that allocates arrays and yields
raw pointers, and a `stream()` callable.  The package instantiates it with torch CUDA tensors
(engine.py); the kernel-logic tests instantiate it with numpy arrays against the simulated
library (tests/sim).  Every method maps 1:1 onto an entry point of include/classpose_b200.h.
"""
from __future__ import annotations

import ctypes as C

from ._abi import Params, check, make_params


class Calls:
    def __init__(self, lib, mem, stream=lambda: None):
        self.lib, self.mem, self.stream = lib, mem, stream

    # ---- helpers -------------------------------------------------------------------------
    def label_capacity(self, H, W):
        return int(self.lib.cpb_label_capacity(H, W))

    def _ws(self, B, H, W, Cc, lcap):
        n = int(self.lib.cpb_workspace_bytes(B, H, W, Cc, lcap))
        ws = self.mem.empty((n,), "uint8")
        return ws, n

    def _p(self, x):
        return None if x is None else self.mem.ptr(x)

    # ---- fused path ------------------------------------------------------------------------
    def compute_masks(self, dP, cellprob, logits=None, params: Params | None = None, want_class_masks=False):
        B, two, H, W = dP.shape
        assert two == 2 and tuple(cellprob.shape) == (B, H, W)
        Cc = 0 if logits is None else int(logits.shape[1])
        prm = params or make_params()
        LC = self.label_capacity(H, W)
        masks = self.mem.empty((B, H, W), "int32")
        counts = self.mem.empty((B,), "int32")
        cell_class = self.mem.zeros((B, LC), "int32") if logits is not None else None
        class_masks = self.mem.empty((B, H, W), "uint8") if (logits is not None and want_class_masks) else None
        ws, n = self._ws(B, H, W, Cc, 0)
        rc = self.lib.cpb_compute_masks_device(self._p(dP), self._p(cellprob), self._p(logits), B, H, W, Cc,
                                               C.byref(prm), self._p(masks), self._p(counts), self._p(cell_class),
                                               self._p(class_masks), self._p(ws), n, self.stream())
        check(rc, "cpb_compute_masks_device")
        self.mem.keep_alive(ws, dP, cellprob, logits)
        return masks, counts, cell_class, class_masks

    def compute_masks_profiled(self, dP, cellprob, logits=None, params: Params | None = None):
        """The fused path once with CUDA events around every stage (synchronises the stream).
        Returns (masks, counts, cell_class, {stage: ms}, flow-check counters)."""
        B, two, H, W = dP.shape
        Cc = 0 if logits is None else int(logits.shape[1])
        prm = params or make_params()
        LC = self.label_capacity(H, W)
        masks = self.mem.empty((B, H, W), "int32")
        counts = self.mem.empty((B,), "int32")
        cell_class = self.mem.zeros((B, LC), "int32") if logits is not None else None
        ws, n = self._ws(B, H, W, Cc, 0)
        ns = int(self.lib.cpb_num_stages())
        ms = (C.c_float * ns)()
        rc = self.lib.cpb_compute_masks_profiled_device(self._p(dP), self._p(cellprob), self._p(logits), B, H, W, Cc,
                                                        C.byref(prm), self._p(masks), self._p(counts),
                                                        self._p(cell_class), None, self._p(ws), n, self.stream(), ms)
        check(rc, "cpb_compute_masks_profiled_device")
        qc = (C.c_int32 * 24)()
        self.lib.cpb_debug_qc_stats(qc)
        stats = {"screen_jobs": int(qc[0]), "float64_labels": int(qc[2]), "screen_decided": int(qc[4]),
                 "screen_undecided": int(qc[5])}
        stages = {self.lib.cpb_stage_name(i).decode(): float(ms[i]) for i in range(ns)}
        return masks, counts, cell_class, stages, stats

    # ---- stages ----------------------------------------------------------------------------
    def follow_flows(self, dP, cellprob, niter=200, cellprob_threshold=0.0, want_float=False):
        B, _, H, W = dP.shape
        p_final = self.mem.empty((B, H, W), "int32")
        p_float = self.mem.zeros((B, 2, H, W), "float32") if want_float else None
        ws, n = self._ws(B, H, W, 0, 0)
        rc = self.lib.cpb_follow_flows_device(self._p(dP), self._p(cellprob), B, H, W, int(niter),
                                              float(cellprob_threshold), self._p(p_final), self._p(p_float),
                                              self._p(ws), n, self.stream())
        check(rc, "cpb_follow_flows_device")
        self.mem.keep_alive(ws, dP, cellprob)
        return p_final, p_float

    def get_masks(self, p_final, max_size_fraction=0.4):
        B, H, W = p_final.shape
        masks = self.mem.empty((B, H, W), "int32")
        counts = self.mem.empty((B,), "int32")
        ws, n = self._ws(B, H, W, 0, 0)
        rc = self.lib.cpb_get_masks_device(self._p(p_final), B, H, W, float(max_size_fraction), self._p(masks),
                                           self._p(counts), self._p(ws), n, self.stream())
        check(rc, "cpb_get_masks_device")
        self.mem.keep_alive(ws, p_final)
        return masks, counts

    def masks_to_flows(self, masks, lcap):
        B, H, W = masks.shape
        mu = self.mem.empty((B, 2, H, W), "float64")
        ws, n = self._ws(B, H, W, 0, lcap)
        rc = self.lib.cpb_masks_to_flows_device(self._p(masks), B, H, W, int(lcap), self._p(mu), self._p(ws), n,
                                                self.stream())
        check(rc, "cpb_masks_to_flows_device")
        self.mem.keep_alive(ws, masks)
        return mu

    def remove_bad_flow_masks(self, masks, dP, lcap, threshold=0.4, want_err=False):
        """In place on `masks`; returns (masks, flow_err or None)."""
        B, H, W = masks.shape
        err = self.mem.zeros((B, lcap), "float64") if want_err else None
        ws, n = self._ws(B, H, W, 0, lcap)
        rc = self.lib.cpb_remove_bad_flow_masks_device(self._p(masks), self._p(dP), B, H, W, int(lcap),
                                                       float(threshold), self._p(err), self._p(ws), n, self.stream())
        check(rc, "cpb_remove_bad_flow_masks_device")
        self.mem.keep_alive(ws, masks, dP)
        return masks, err

    def fill_holes_and_remove_small_masks(self, masks, lcap, min_size=15):
        B, H, W = masks.shape
        counts = self.mem.empty((B,), "int32")
        ws, n = self._ws(B, H, W, 0, lcap)
        rc = self.lib.cpb_fill_holes_and_remove_small_masks_device(self._p(masks), B, H, W, int(lcap), int(min_size),
                                                                   self._p(counts), self._p(ws), n, self.stream())
        check(rc, "cpb_fill_holes_and_remove_small_masks_device")
        self.mem.keep_alive(ws, masks)
        return masks, counts

    def class_vote(self, masks, logits, lcap, want_class_masks=True):
        B, H, W = masks.shape
        Cc = int(logits.shape[1])
        cell_class = self.mem.zeros((B, lcap), "int32")
        class_masks = self.mem.empty((B, H, W), "uint8") if want_class_masks else None
        ws, n = self._ws(B, H, W, Cc, lcap)
        rc = self.lib.cpb_class_vote_device(self._p(masks), self._p(logits), B, H, W, Cc, int(lcap),
                                            self._p(cell_class), self._p(class_masks), self._p(ws), n, self.stream())
        check(rc, "cpb_class_vote_device")
        self.mem.keep_alive(ws, masks, logits)
        return cell_class, class_masks

    def remove_border_instances(self, masks, lcap, nch=1):
        B, H, W = masks.shape[:3]
        ws, n = self._ws(B, H, W, 0, lcap)
        rc = self.lib.cpb_remove_border_instances_device(self._p(masks), B, H, W, int(nch), int(lcap), self._p(ws), n,
                                                         self.stream())
        check(rc, "cpb_remove_border_instances_device")
        self.mem.keep_alive(ws, masks)
        return masks

    def average_tiles(self, y, y0, x0, flip, negate_flow, taper_y, taper_x, Ly, Lx, crop=(0, 0, 0, 0),
                      x0_multiple_of_4=False, max_cover=0):
        """x0_multiple_of_4 / max_cover: host-side knowledge of the window geometry (see tile_cover) that
        enables the 128-bit kernel; leave at the defaults when unknown."""
        B, ntiles, nch, ly, lx = y.shape
        cy0, cy1, cx0, cx1 = crop
        yf = self.mem.empty((B, nch, Ly - cy0 - cy1, Lx - cx0 - cx1), "float32")
        rc = self.lib.cpb_average_tiles_ex_device(self._p(y), B, ntiles, nch, ly, lx, self._p(y0), self._p(x0),
                                                  self._p(flip), 1 if negate_flow else 0, self._p(taper_y),
                                                  self._p(taper_x), int(Ly), int(Lx), cy0, cy1, cx0, cx1, self._p(yf),
                                                  1 if x0_multiple_of_4 else 0, int(max_cover), self.stream())
        check(rc, "cpb_average_tiles_ex_device")
        self.mem.keep_alive(y, y0, x0, flip, taper_y, taper_x)
        return yf

    def eval_tail(self, y_flows, y_logits, y0, x0, flip, augment, taper_y, taper_x, Ly, Lx, crop, params: Params | None = None,
                  want_class_masks=False):
        """Fused tail of ClassposeModel.eval on the network's sub-tile outputs (cpb_eval_tail_device).
        Returns (masks, counts, cell_class, class_masks, dP, cellprob, logits)."""
        B, nt, three, ly, lx = y_flows.shape
        assert three == 3
        Cc = 0 if y_logits is None else int(y_logits.shape[2])
        cy0, cy1, cx0, cx1 = crop
        H, W = Ly - cy0 - cy1, Lx - cx0 - cx1
        prm = params or make_params()
        LC = self.label_capacity(H, W)
        dP = self.mem.empty((B, 2, H, W), "float32")
        cellprob = self.mem.empty((B, H, W), "float32")
        logits = self.mem.empty((B, Cc, H, W), "float32") if Cc else None
        masks = self.mem.empty((B, H, W), "int32")
        counts = self.mem.empty((B,), "int32")
        cell_class = self.mem.zeros((B, LC), "int32") if Cc else None
        class_masks = self.mem.empty((B, H, W), "uint8") if (Cc and want_class_masks) else None
        ws, n = self._ws(B, H, W, Cc, 0)
        rc = self.lib.cpb_eval_tail_device(self._p(y_flows), self._p(y_logits), B, nt, Cc, ly, lx, self._p(y0), self._p(x0),
                                           self._p(flip), 1 if augment else 0, self._p(taper_y), self._p(taper_x), int(Ly),
                                           int(Lx), cy0, cy1, cx0, cx1, C.byref(prm), self._p(dP), self._p(cellprob),
                                           self._p(logits), self._p(masks), self._p(counts), self._p(cell_class),
                                           self._p(class_masks), self._p(ws), n, self.stream())
        check(rc, "cpb_eval_tail_device")
        self.mem.keep_alive(ws, y_flows, y_logits, y0, x0, flip, taper_y, taper_x)
        return masks, counts, cell_class, class_masks, dP, cellprob, logits

    def cell_contours(self, masks, lcap, points_cap=None):
        """PostProcessor features per label: dict(npoints, offsets, total, points, feat, perimeter, valid)."""
        B, H, W = masks.shape
        cap = int(points_cap) if points_cap else max(1024, B * H * W // 8)
        out = dict(npoints=self.mem.zeros((B, lcap), "int32"), offsets=self.mem.zeros((B, lcap), "int64"),
                   total=self.mem.zeros((1,), "int64"), points=self.mem.empty((cap, 2), "int16"),
                   feat=self.mem.zeros((B, lcap, 8), "int64"), perimeter=self.mem.zeros((B, lcap), "float64"),
                   valid=self.mem.zeros((B, lcap), "int32"))
        ws, n = self._ws(B, H, W, 0, lcap)
        rc = self.lib.cpb_cell_contours_device(self._p(masks), B, H, W, int(lcap), self._p(out["npoints"]),
                                               self._p(out["offsets"]), self._p(out["total"]), self._p(out["points"]), cap,
                                               self._p(out["feat"]), self._p(out["perimeter"]), self._p(out["valid"]),
                                               self._p(ws), n, self.stream())
        check(rc, "cpb_cell_contours_device")
        self.mem.keep_alive(ws, masks)
        return out

    def dedup_cells(self, cx, cy, size, max_dist=7.5, want_group=False):
        n = int(cx.shape[0])
        keep = self.mem.zeros((max(n, 1),), "int32")
        group = self.mem.zeros((max(n, 1),), "int32") if want_group else None
        nb = int(self.lib.cpb_dedup_workspace_bytes(n))
        ws = self.mem.empty((nb,), "uint8")
        rc = self.lib.cpb_dedup_cells_device(self._p(cx), self._p(cy), self._p(size), n, float(max_dist), self._p(keep),
                                             self._p(group), self._p(ws), nb, self.stream())
        check(rc, "cpb_dedup_cells_device")
        self.mem.keep_alive(ws, cx, cy, size)
        return keep[:n], (group[:n] if want_group else None)

    def prepare_tiles(self, img, pads, y0, x0, flip, ly, lx, lower=1.0, upper=99.0):
        """img [B,H,W,C] float32 -> (tiles [B,ntiles,C,ly,lx], lowhigh [B,C,2], code [B,C])."""
        B, H, W, Cc = img.shape
        nt = int(y0.shape[0])
        tiles = self.mem.empty((B, nt, Cc, ly, lx), "float32")
        lowhigh = self.mem.empty((B, Cc, 2), "float32")
        code = self.mem.empty((B, Cc), "int32")
        rc = self.lib.cpb_prepare_tiles_device(self._p(img), B, H, W, Cc, float(lower), float(upper), int(pads[0]),
                                               int(pads[2]), nt, int(ly), int(lx), self._p(y0), self._p(x0), self._p(flip),
                                               self._p(tiles), self._p(lowhigh), self._p(code), self.stream())
        check(rc, "cpb_prepare_tiles_device")
        self.mem.keep_alive(img, y0, x0, flip)
        return tiles, lowhigh, code

    def label_offsets(self, counts, base=0):
        B = counts.shape[0]
        offsets = self.mem.empty((B,), "int64")
        total = self.mem.empty((1,), "int64")
        rc = self.lib.cpb_label_offsets_device(self._p(counts), B, int(base), self._p(offsets), self._p(total),
                                               self.stream())
        check(rc, "cpb_label_offsets_device")
        self.mem.keep_alive(counts)
        return offsets, total
