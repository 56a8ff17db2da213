"""ctypes declaration of the C ABI in include/classpose_b200.h (signatures only, no logic)."""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 2

E_ARG, E_WORKSPACE, E_RANGE, E_CAPACITY = -1, -2, -3, -4
_ERR = {E_ARG: "bad argument (shape / null pointer)", E_WORKSPACE: "workspace too small",
        E_RANGE: "B*H*W exceeds the 31-bit pixel index; split the batch",
        E_CAPACITY: "a tile exhausted an internal pool (counts[b] = -1): its masks are incomplete"}


class Params(C.Structure):
    """cpb_params -- the arguments of dynamics.resize_and_compute_masks as Classpose passes them
    (/root/reference/src/classpose/models.py:149-159; defaults models.py:490-498, 751-752)."""
    _fields_ = [("niter", C.c_int32), ("cellprob_threshold", C.c_float), ("flow_threshold", C.c_double),
                ("min_size", C.c_int32), ("max_size_fraction", C.c_double), ("remove_border", C.c_int32),
                ("fill_holes", C.c_int32)]


class HostOptions(C.Structure):
    """cpb_host_options (include/classpose_b200.h)."""
    _fields_ = [("tiles_per_chunk", C.c_int32), ("device", C.c_int32), ("logits_mode", C.c_int32),
                ("flows_mode", C.c_int32), ("masks_u16", C.c_int32)]


LOGITS_AUTO, LOGITS_UPLOAD, LOGITS_MAPPED = 0, 1, 2


class ClassposeB200Error(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise ClassposeB200Error(f"{what}: {_ERR.get(rc, 'error %d' % rc)}")
    raise ClassposeB200Error(f"{what}: CUDA error {rc}")


_P, _I, _D, _F, _Z, _L = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_size_t, C.c_int64

SIGNATURES = {
    "cpb_abi_version": (C.c_int, []),
    "cpb_label_capacity": (C.c_int, [_I, _I]),
    "cpb_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "cpb_compute_masks_device": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(Params), _P, _P, _P, _P, _P, _Z, _P]),
    "cpb_num_stages": (C.c_int, []),
    "cpb_stage_name": (C.c_char_p, [_I]),
    "cpb_compute_masks_profiled_device": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(Params), _P, _P, _P, _P, _P, _Z, _P, _P]),
    "cpb_debug_launch_count": (C.c_longlong, []),
    "cpb_debug_qc_stats": (None, [_P]),
    "cpb_debug_set_follow_merge": (None, [_I]),
    "cpb_debug_set_switch": (None, [_I, _I]),
    "cpb_compute_masks_host": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(Params), _P, _P, _P, _P, _I, _I]),
    "cpb_compute_masks_host_ex": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(Params), _P, _P, _P, _P, C.POINTER(HostOptions)]),
    "cpb_tile_plan_create": (C.c_int, [_I, _I, C.POINTER(Params), _I, C.POINTER(C.c_void_p)]),
    "cpb_tile_plan_destroy": (None, [_P]),
    "cpb_tile_plan_dp": (_P, [_P]),
    "cpb_tile_plan_cellprob": (_P, [_P]),
    "cpb_tile_plan_masks": (_P, [_P]),
    "cpb_tile_plan_run": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "cpb_tile_plan_logits": (_P, [_P, _I]),
    "cpb_tile_plan_vote": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "cpb_follow_flows_device": (C.c_int, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _Z, _P]),
    "cpb_get_masks_device": (C.c_int, [_P, _I, _I, _I, _D, _P, _P, _P, _Z, _P]),
    "cpb_masks_to_flows_device": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "cpb_remove_bad_flow_masks_device": (C.c_int, [_P, _P, _I, _I, _I, _I, _D, _P, _P, _Z, _P]),
    "cpb_fill_holes_and_remove_small_masks_device": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "cpb_class_vote_device": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "cpb_class_vote_counts_device": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "cpb_remove_border_instances_device": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "cpb_average_tiles_device": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "cpb_average_tiles_ex_device": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "cpb_eval_tail_device": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I,
                                       C.POINTER(Params), _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "cpb_cell_contours_device": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _L, _P, _P, _P, _P, _Z, _P]),
    "cpb_dedup_workspace_bytes": (_Z, [_L]),
    "cpb_dedup_cells_device": (C.c_int, [_P, _P, _P, _L, _D, _P, _P, _P, _Z, _P]),
    "cpb_prepare_tiles_device": (C.c_int, [_P, _I, _I, _I, _I, _D, _D, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "cpb_label_offsets_device": (C.c_int, [_P, _I, _L, _P, _P, _P]),
}

# entry points that exist only in the CUDA build (host-buffer path does real H2D/D2H copies)
CUDA_ONLY = {"cpb_compute_masks_host", "cpb_compute_masks_host_ex", "cpb_tile_plan_create", "cpb_tile_plan_destroy",
             "cpb_tile_plan_dp", "cpb_tile_plan_cellprob", "cpb_tile_plan_masks", "cpb_tile_plan_run", "cpb_tile_plan_logits",
             "cpb_tile_plan_vote"}


def declare(lib: C.CDLL, cuda: bool = True) -> C.CDLL:
    """Attach argtypes/restype for every exported symbol; raises if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        if not cuda and name in CUDA_ONLY:
            continue
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ClassposeB200Error(f"library does not export {name}") from e
        fn.restype, fn.argtypes = res, args
    if lib.cpb_abi_version() != ABI_VERSION:
        raise ClassposeB200Error("ABI version mismatch between the python host and the shared library")
    return lib


def make_params(niter=200, cellprob_threshold=0.0, flow_threshold=0.4, min_size=15, max_size_fraction=0.4,
                remove_border=False, fill_holes=True) -> Params:
    ft = 0.0 if flow_threshold is None else float(flow_threshold)
    return Params(int(niter), float(cellprob_threshold), ft, int(min_size), float(max_size_fraction),
                  1 if remove_border else 0, 1 if fill_holes else 0)
