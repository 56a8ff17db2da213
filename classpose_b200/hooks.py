"""Attribute replacement of the four hooks the reference resolves at call time:

  A  classpose.models.compute_masks                 (models.py:464 -> :97)
  B  cellpose.dynamics.resize_and_compute_masks / compute_masks   (models.py:120, 149)
  C  classpose.models.compute_class_masks           (models.py:766 -> :191)
  D  cellpose.transforms.average_tiles              (core.py:215, 218)
  D' cellpose.transforms.unaugment_tiles            (core.py:209, --tta) and classpose.core.unaugment_class_tiles
     (core.py:8 binds the NAME at import time, so the attribute of classpose.core is what core.py:213 resolves;
     the defining modules classpose.transforms[.transforms] are patched too for later importers)

`install()` patches whichever of those modules can be imported in the running interpreter and
returns the list of patched names; `uninstall()` restores the originals.
"""
from __future__ import annotations

import importlib

from . import dynamics as _dyn
from . import models as _models
from . import transforms as _tf
from . import utils as _utils

_saved = {}

_HOOKS = [
    ("classpose.models", "compute_masks", _models.compute_masks),
    ("classpose.models", "compute_class_masks", _models.compute_class_masks),
    ("cellpose.dynamics", "resize_and_compute_masks", _dyn.resize_and_compute_masks),
    ("cellpose.dynamics", "compute_masks", _dyn.compute_masks),
    ("cellpose.utils", "fill_holes_and_remove_small_masks", _utils.fill_holes_and_remove_small_masks),
    ("cellpose.transforms", "average_tiles", _tf.average_tiles),
    ("cellpose.transforms", "unaugment_tiles", _tf.unaugment_tiles),
    ("classpose.core", "unaugment_class_tiles", _tf.unaugment_class_tiles),
    ("classpose.transforms", "unaugment_class_tiles", _tf.unaugment_class_tiles),
    ("classpose.transforms.transforms", "unaugment_class_tiles", _tf.unaugment_class_tiles),
]


def install(modules=None):
    """modules: optional dict name -> module object to patch instead of importing (used by tests)."""
    patched = []
    for modname, attr, fn in _HOOKS:
        try:
            mod = modules[modname] if modules and modname in modules else importlib.import_module(modname)
        except Exception:
            continue
        key = (modname, attr)
        if key not in _saved:
            _saved[key] = (mod, getattr(mod, attr, None))
        setattr(mod, attr, fn)
        patched.append(f"{modname}.{attr}")
    return patched


def uninstall():
    for (modname, attr), (mod, orig) in list(_saved.items()):
        if orig is None:
            try:
                delattr(mod, attr)
            except AttributeError:
                pass
        else:
            setattr(mod, attr, orig)
        del _saved[(modname, attr)]
