"""classpose_b200 -- B200-native (sm_100a) implementation of Classpose's per-tile post-network
path: network flows / cell probability / class logits -> instance masks -> per-cell class.

Layout: csrc/ holds the CUDA kernels and the C ABI (include/classpose_b200.h); the python
modules mirror the reference's call boundary (dynamics, utils, transforms, models, metrics)
and add the batched device API (engine) and the multi-GPU label offsets (distributed).
There is no CPU fallback: without a CUDA device or the built library every compute call raises.
"""
from ._abi import ClassposeB200Error, make_params  # noqa: F401

__version__ = "0.1.0"

__all__ = ["ClassposeB200Error", "make_params", "get_engine", "Engine", "install", "uninstall",
           "compute_masks", "compute_class_masks", "resize_and_compute_masks", "average_tiles",
           "remove_border_instances", "fill_holes_and_remove_small_masks"]


def __getattr__(name):  # lazy: importing the package must not require torch/CUDA to be usable
    if name in ("get_engine", "Engine"):
        from . import engine
        return getattr(engine, name)
    if name in ("install", "uninstall"):
        import importlib
        return getattr(importlib.import_module(__name__ + ".hooks"), name)
    if name in ("compute_masks", "compute_class_masks"):
        from . import models
        return getattr(models, name)
    if name == "resize_and_compute_masks":
        from . import dynamics
        return dynamics.resize_and_compute_masks
    if name == "average_tiles":
        from . import transforms
        return transforms.average_tiles
    if name == "remove_border_instances":
        from . import metrics
        return metrics.remove_border_instances
    if name == "fill_holes_and_remove_small_masks":
        from . import utils
        return utils.fill_holes_and_remove_small_masks
    raise AttributeError(name)
