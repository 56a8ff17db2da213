"""Device-resident batched API: torch CUDA tensors in, torch CUDA tensors out.

torch is plumbing here (device memory, current stream); all arithmetic runs in the
hand-written sm_100a kernels of libclasspose_b200.so.  Calls are asynchronous on the
current torch stream and re-entrant (each call allocates its own workspace from torch's
stream-ordered caching allocator), so the reference's two inference threads per process
(/root/reference/src/classpose/entrypoints/predict_wsi.py:728-797) can share one Engine.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._abi import LOGITS_AUTO, LOGITS_MAPPED, LOGITS_UPLOAD, ClassposeB200Error, HostOptions, check, make_params
from ._calls import Calls

_DT = {"int32": torch.int32, "int16": torch.int16, "uint8": torch.uint8, "float32": torch.float32, "float64": torch.float64,
       "int64": torch.int64}


class _TorchMem:
    def __init__(self, device):
        self.device = device

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=_DT[dtype], device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=_DT[dtype], device=self.device)

    def ptr(self, x):
        if not x.is_cuda or not x.is_contiguous():
            raise ClassposeB200Error("expected a contiguous CUDA tensor")
        return x.data_ptr()

    def keep_alive(self, *a):
        pass


class Engine:
    """One per (process, device).  Thread-safe."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise ClassposeB200Error("CUDA device required: classpose_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ClassposeB200Error(f"device {self.device} is not a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.calls = Calls(self.lib, _TorchMem(self.device),
                           stream=lambda: torch.cuda.current_stream(self.device).cuda_stream)

    # -- helpers ---------------------------------------------------------------------------
    def _dev(self, x, dtype):
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        return x.to(self.device, dtype=dtype, non_blocking=True).contiguous()

    def label_capacity(self, H, W):
        return self.calls.label_capacity(H, W)

    # -- fused path ------------------------------------------------------------------------
    MAX_PIXELS_PER_CALL = 2 ** 31 - 2 ** 24      # one device call indexes pixels with 31 bits (CPB_E_RANGE)

    def compute_masks_batch(self, dP, cellprob, logits=None, niter=200, cellprob_threshold=0.0,
                            flow_threshold=0.4, min_size=15, max_size_fraction=0.4, remove_border=False,
                            fill_holes=True, want_class_masks=False, max_pixels_per_call=None):
        """dP [B,2,H,W], cellprob [B,H,W], logits [B,C,H,W] (float32, CUDA or numpy)
        -> masks int32 [B,H,W], counts int32 [B], cell_class int32 [B,LC] | None, class_masks uint8 | None.
        Batches beyond the 31-bit pixel index of one device call are processed in slices."""
        with torch.cuda.device(self.device):
            dP = self._dev(dP, torch.float32)
            cellprob = self._dev(cellprob, torch.float32)
            logits = self._dev(logits, torch.float32)
            prm = make_params(niter, cellprob_threshold, flow_threshold, min_size, max_size_fraction,
                              remove_border, fill_holes)
            B, _, H, W = dP.shape
            per = max(1, (max_pixels_per_call or self.MAX_PIXELS_PER_CALL) // (H * W))
            if B <= per:
                return self.calls.compute_masks(dP, cellprob, logits, prm, want_class_masks)
            parts = [self.calls.compute_masks(dP[i:i + per], cellprob[i:i + per],
                                              None if logits is None else logits[i:i + per], prm, want_class_masks)
                     for i in range(0, B, per)]
            return tuple(None if parts[0][k] is None else torch.cat([p[k] for p in parts]) for k in range(4))

    def compute_masks_host(self, dP, cellprob, logits=None, niter=200, cellprob_threshold=0.0, flow_threshold=0.4,
                           min_size=15, max_size_fraction=0.4, remove_border=False, fill_holes=True,
                           want_class_masks=False, tiles_per_chunk=0, out=None, logits_mode="auto", flows_mode="auto",
                           masks_u16=False):
        """Host buffers in / out (numpy arrays or CPU torch tensors, ideally pinned).  The library
        performs chunked H2D -> kernels -> D2H itself.  Returns (masks, counts, cell_class, class_masks) as CPU
        tensors.  `out` may hold pre-allocated (pinned) output tensors with the same keys.
        logits_mode / flows_mode: "auto" (a pinned buffer is read in place through its mapped pointer -- logits
        only under cells, dP only where there is foreground), "upload" (copy everything) or "mapped" (insist);
        masks_u16: deliver uint16 label images."""
        def host(x, dtype):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(np.ascontiguousarray(x, dtype=dtype))
            if x.is_cuda or not x.is_contiguous() or x.dtype != torch.float32:
                raise ClassposeB200Error("compute_masks_host expects contiguous float32 host buffers")
            return x
        dP, cellprob, logits = host(dP, np.float32), host(cellprob, np.float32), host(logits, np.float32)
        B, _, H, W = dP.shape
        Cc = 0 if logits is None else int(logits.shape[1])
        LC = self.label_capacity(H, W)
        out = out or {}
        mdt = torch.uint16 if masks_u16 else torch.int32
        masks = out.get("masks") if out.get("masks") is not None else torch.empty((B, H, W), dtype=mdt)
        if masks.dtype != mdt:
            raise ClassposeB200Error(f"out['masks'] must be {mdt}")
        counts = out.get("counts") if out.get("counts") is not None else torch.empty((B,), dtype=torch.int32)
        cell_class = None
        class_masks = None
        if logits is not None:
            cell_class = out.get("cell_class") if out.get("cell_class") is not None else torch.zeros((B, LC), dtype=torch.int32)
            if want_class_masks:
                class_masks = out.get("class_masks") if out.get("class_masks") is not None else torch.empty((B, H, W), dtype=torch.uint8)
        prm = make_params(niter, cellprob_threshold, flow_threshold, min_size, max_size_fraction, remove_border,
                          fill_holes)
        p = lambda t: None if t is None else t.data_ptr()
        modes = {"auto": LOGITS_AUTO, "upload": LOGITS_UPLOAD, "mapped": LOGITS_MAPPED}
        opt = HostOptions(int(tiles_per_chunk), int(self.device.index), modes[logits_mode], modes[flows_mode],
                          1 if masks_u16 else 0)
        rc = self.lib.cpb_compute_masks_host_ex(p(dP), p(cellprob), p(logits), B, H, W, Cc, C.byref(prm), p(masks),
                                                p(counts), p(cell_class), p(class_masks), C.byref(opt))
        check(rc, "cpb_compute_masks_host_ex")
        return masks, counts, cell_class, class_masks

    # -- stages (device tensors) -----------------------------------------------------------
    def follow_flows(self, dP, cellprob, niter=200, cellprob_threshold=0.0, want_float=False):
        with torch.cuda.device(self.device):
            return self.calls.follow_flows(self._dev(dP, torch.float32), self._dev(cellprob, torch.float32), niter,
                                           cellprob_threshold, want_float)

    def get_masks(self, p_final, max_size_fraction=0.4):
        with torch.cuda.device(self.device):
            return self.calls.get_masks(self._dev(p_final, torch.int32), max_size_fraction)

    def masks_to_flows(self, masks, lcap):
        with torch.cuda.device(self.device):
            return self.calls.masks_to_flows(self._dev(masks, torch.int32), lcap)

    def remove_bad_flow_masks(self, masks, dP, lcap, threshold=0.4, want_err=False):
        with torch.cuda.device(self.device):
            return self.calls.remove_bad_flow_masks(self._dev(masks, torch.int32).clone(), self._dev(dP, torch.float32),
                                                    lcap, threshold, want_err)

    def fill_holes_and_remove_small_masks(self, masks, lcap, min_size=15):
        with torch.cuda.device(self.device):
            return self.calls.fill_holes_and_remove_small_masks(self._dev(masks, torch.int32).clone(), lcap, min_size)

    def class_vote(self, masks, logits, lcap, want_class_masks=True):
        with torch.cuda.device(self.device):
            return self.calls.class_vote(self._dev(masks, torch.int32), self._dev(logits, torch.float32), lcap,
                                         want_class_masks)

    def remove_border_instances(self, masks, lcap, nch=1):
        with torch.cuda.device(self.device):
            return self.calls.remove_border_instances(self._dev(masks, torch.int32).clone(), lcap, nch)

    def average_tiles(self, y, y0, x0, flip, negate_flow, taper_y, taper_x, Ly, Lx, crop=(0, 0, 0, 0),
                      x0_multiple_of_4=False, max_cover=0):
        with torch.cuda.device(self.device):
            return self.calls.average_tiles(self._dev(y, torch.float32), self._dev(y0, torch.int32),
                                            self._dev(x0, torch.int32), self._dev(flip, torch.int32), negate_flow,
                                            self._dev(taper_y, torch.float64), self._dev(taper_x, torch.float64),
                                            Ly, Lx, crop, x0_multiple_of_4, max_cover)

    def cell_contours(self, masks, lcap, points_cap=None):
        with torch.cuda.device(self.device):
            return self.calls.cell_contours(self._dev(masks, torch.int32), lcap, points_cap)

    def dedup_cells(self, cx, cy, size, max_dist=7.5, want_group=False):
        with torch.cuda.device(self.device):
            return self.calls.dedup_cells(self._dev(cx, torch.float64), self._dev(cy, torch.float64),
                                          self._dev(size, torch.float64), max_dist, want_group)

    def prepare_tiles(self, img, pads, y0, x0, flip, ly, lx, lower=1.0, upper=99.0):
        with torch.cuda.device(self.device):
            return self.calls.prepare_tiles(self._dev(img, torch.float32), pads, self._dev(y0, torch.int32),
                                            self._dev(x0, torch.int32), self._dev(flip, torch.int32), ly, lx, lower, upper)

    def label_offsets(self, counts, base=0):
        with torch.cuda.device(self.device):
            return self.calls.label_offsets(self._dev(counts, torch.int32), base)


_engines = {}


def get_engine(device=None) -> Engine:
    """Process-wide engine cache keyed by device."""
    if not torch.cuda.is_available():
        raise ClassposeB200Error("CUDA device required: classpose_b200 has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        # the reference passes the model's device through; a CPU device means "use the GPU this
        # process owns" here -- inputs are uploaded, never computed on the host
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = dev.index
    if key not in _engines:
        _engines[key] = Engine(dev)
    return _engines[key]


def _profile_stages(self, dP, cellprob, logits=None, with_qc=False, **kw):
    """Fused path once with CUDA events around every stage -> {stage name: device ms}.  Synchronises.
    with_qc: also return the flow-check counters (float32 screen decided / undecided, float64 labels)."""
    with torch.cuda.device(self.device):
        _, _, _, stages, qc = self.calls.compute_masks_profiled(dP, cellprob, logits, make_params(**kw))
        return (stages, qc) if with_qc else stages


Engine.profile_stages = _profile_stages
Engine.launch_count = lambda self: int(self.lib.cpb_debug_launch_count())
