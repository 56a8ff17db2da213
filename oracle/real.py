"""Locate the REAL implementation of this path -- the `cellpose` package the reference imports
(cellpose==4.0.8, /root/reference/uv.lock:352-353) -- if it can be imported in the running interpreter.

TEST INFRASTRUCTURE (see oracle/__init__.py).  It is absent from this image and from the GPU box today, so
everything that uses this module degrades to the oracle port; the day `pip install cellpose` (or a
`baseline/_ref` install of the reference) exists, the same tests and the same bench legs run against the real
thing with no code change:

  * tests/test_real_cellpose.py   diffs oracle/ against it on the committed fixtures (pins the port),
  * bench.py --impl reference / cpu_baseline   time it instead of the port (`"kind": "reference"`).
"""
from __future__ import annotations

import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def find():
    """-> dict(kind='reference', dynamics=..., utils=..., transforms=..., version=..., where=...) or None."""
    tried = []
    for extra in (None, _REF_DIR):
        if extra is not None:
            if not os.path.isdir(extra):
                continue
            if extra not in sys.path:
                sys.path.insert(0, extra)
        try:
            cp = importlib.import_module("cellpose")
            dyn = importlib.import_module("cellpose.dynamics")
            utl = importlib.import_module("cellpose.utils")
            tfm = importlib.import_module("cellpose.transforms")
        except Exception as e:          # ModuleNotFoundError today; a broken install must not break the tests either
            tried.append(f"{extra or 'sys.path'}: {type(e).__name__}: {e}")
            continue
        need = [(dyn, "resize_and_compute_masks"), (dyn, "compute_masks"), (dyn, "follow_flows"),
                (dyn, "remove_bad_flow_masks"), (dyn, "masks_to_flows"), (utl, "fill_holes_and_remove_small_masks"),
                (tfm, "average_tiles")]
        missing = [f"{m.__name__}.{a}" for m, a in need if not hasattr(m, a)]
        if missing:
            tried.append(f"{extra or 'sys.path'}: cellpose found but lacks {missing}")
            continue
        version = getattr(cp, "version", None) or getattr(cp, "__version__", "unknown")
        return dict(kind="reference", dynamics=dyn, utils=utl, transforms=tfm, version=str(version),
                    where=os.path.dirname(cp.__file__))
    find.tried = tried
    return None


find.tried = []


def resize_and_compute_masks(real, dP, cellprob, **kw):
    """The reference's own call (models.py:149-159) on the real package, CPU device."""
    import torch
    return real["dynamics"].resize_and_compute_masks(dP, cellprob, device=torch.device("cpu"), **kw)
