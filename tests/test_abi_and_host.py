"""CPU-side checks: the CUDA library loads and exports every symbol the header declares, the host
logic mirrors the reference's geometry, the hooks patch by attribute, and there is no CPU fallback."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

import classpose_b200
from classpose_b200 import _abi, _lib, distributed as cdist, hooks, transforms as btf
from oracle import transforms as otf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    path = _lib.build()
    lib = ctypes.CDLL(path)          # no GPU needed to load
    header = open(os.path.join(ROOT, "include", "classpose_b200.h")).read()
    names = set(re.findall(r"\b(cpb_[a-z_0-9]+)\s*\(", header))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert names == set(_abi.SIGNATURES), names ^ set(_abi.SIGNATURES)
    _abi.declare(lib)
    assert lib.cpb_abi_version() == _abi.ABI_VERSION
    assert lib.cpb_label_capacity(256, 256) == 256 * 256 // 11 + 2
    assert lib.cpb_workspace_bytes(4, 256, 256, 7, 0) > 4 * 256 * 256 * 36


def test_params_struct_layout_matches_header():
    p = _abi.make_params()
    assert (p.niter, p.min_size, p.fill_holes, p.remove_border) == (200, 15, 1, 0)
    assert abs(p.flow_threshold - 0.4) < 1e-15 and abs(p.max_size_fraction - 0.4) < 1e-15
    assert ctypes.sizeof(_abi.Params) == 40   # int,float,double,int,(pad),double,int,int; static_assert in cpb_api.cu
    assert _abi.make_params(flow_threshold=None).flow_threshold == 0.0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the GPU-less behaviour")
def test_no_cpu_fallback():
    with pytest.raises(classpose_b200.ClassposeB200Error):
        classpose_b200.get_engine()
    from classpose_b200 import dynamics
    with pytest.raises(classpose_b200.ClassposeB200Error):
        dynamics.resize_and_compute_masks(np.zeros((2, 8, 8), np.float32), np.ones((8, 8), np.float32))
    with pytest.raises(classpose_b200.ClassposeB200Error):
        classpose_b200.compute_class_masks(np.ones((4, 4), np.int32), np.zeros((3, 1, 4, 4), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "classpose_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "cusim" not in src or f == "cpb_platform.h", f


@pytest.mark.parametrize("Ly,Lx,bsize,augment", [(272, 272, 256, False), (272, 272, 256, True), (528, 400, 256, False),
                                                  (300, 1040, 256, True), (256, 256, 256, False), (200, 180, 224, True)])
def test_tile_geometry_matches_make_tiles(Ly, Lx, bsize, augment):
    img = np.zeros((1, Ly, Lx), np.float32)
    IMG, ysub, xsub, Lyt, Lxt = otf.make_tiles(img, bsize=bsize, augment=augment, tile_overlap=0.1)
    g = btf.tile_geometry(Ly, Lx, bsize, augment=augment, tile_overlap=0.1)
    assert (g["Ly"], g["Lx"]) == (Lyt, Lxt)
    assert list(g["y0"]) == [s[0] for s in ysub] and list(g["x0"]) == [s[0] for s in xsub]
    assert (g["ly"], g["lx"]) == IMG.shape[-2:]
    if augment:
        codes = [otf._flip_code(j, i) for j in range(g["ny"]) for i in range(g["nx"])]
        assert list(g["flip"]) == codes


def test_taper_and_pad_match_oracle():
    for ly, lx in ((256, 256), (224, 200), (300, 256)):
        ty, tx = btf.taper_1d(ly, lx)
        np.testing.assert_array_equal(np.outer(ty, tx), otf._taper_mask(ly, lx))
    for L in ((256, 256), (225, 301), (100, 100)):
        assert btf.get_pad_yx(*L, min_size=(256, 256)) == otf.get_pad_yx(*L, min_size=(256, 256))


def test_hooks_patch_and_restore_by_attribute():
    fake_models = types.ModuleType("classpose.models")
    fake_models.compute_masks = "orig_a"
    fake_models.compute_class_masks = "orig_c"
    fake_dyn = types.ModuleType("cellpose.dynamics")
    fake_dyn.resize_and_compute_masks = "orig_b"
    fake_dyn.compute_masks = "orig_b2"
    fake_tf = types.ModuleType("cellpose.transforms")
    fake_tf.average_tiles = "orig_d"
    mods = {"classpose.models": fake_models, "cellpose.dynamics": fake_dyn, "cellpose.transforms": fake_tf}
    patched = hooks.install(mods)
    assert {"classpose.models.compute_masks", "classpose.models.compute_class_masks",
            "cellpose.dynamics.resize_and_compute_masks", "cellpose.transforms.average_tiles"} <= set(patched)
    from classpose_b200 import dynamics, models
    assert fake_models.compute_masks is models.compute_masks
    assert fake_dyn.resize_and_compute_masks is dynamics.resize_and_compute_masks
    hooks.uninstall()
    assert fake_models.compute_masks == "orig_a" and fake_tf.average_tiles == "orig_d"


def test_install_on_a_package_tree_with_the_reference_import_style(tmp_path, monkeypatch):
    """install() against real packages laid out and imported the way the reference does it: `from cellpose import
    dynamics, transforms, utils` (attribute lookups at call time, models.py:13-22, 120, 149), module-global functions called
    by bare name from inside the same module (models.py:464, 766) and a NAME import at module load
    (core.py:8 `from classpose.transforms import unaugment_class_tiles`, used bare at core.py:213)."""
    import sys
    import textwrap
    root = tmp_path / "site"
    files = {
        "cellpose/__init__.py": "",
        "cellpose/dynamics.py": "def resize_and_compute_masks(*a, **k):\n    return 'stock-b'\ndef compute_masks(*a, **k):\n    return 'stock-b2'\n",
        "cellpose/utils.py": "def fill_holes_and_remove_small_masks(*a, **k):\n    return 'stock-u'\n",
        "cellpose/transforms.py": "def average_tiles(*a, **k):\n    return 'stock-d'\ndef unaugment_tiles(*a, **k):\n    return 'stock-ua'\n",
        "classpose/__init__.py": "",
        "classpose/transforms/__init__.py": "from .transforms import unaugment_class_tiles\n",
        "classpose/transforms/transforms.py": "def unaugment_class_tiles(y):\n    return 'stock-uc'\n",
        "classpose/core.py": textwrap.dedent("""
            from cellpose import transforms
            from classpose.transforms import unaugment_class_tiles
            def run_net_tail():
                return transforms.unaugment_tiles, unaugment_class_tiles, transforms.average_tiles
            """),
        "classpose/models.py": textwrap.dedent("""
            from cellpose import dynamics, transforms, utils
            def compute_masks(*a, **k):
                return dynamics.resize_and_compute_masks(*a, **k)
            def compute_class_masks(masks, y_class):
                return 'stock-c'
            class ClassposeModel:
                def _compute_masks(self):
                    return compute_masks            # bare module-global lookup, as models.py:464
                def eval_tail(self):
                    return self._compute_masks(), compute_class_masks, dynamics.resize_and_compute_masks, utils.fill_holes_and_remove_small_masks
            """),
    }
    for rel, src in files.items():
        f = root / rel
        f.parent.mkdir(parents=True, exist_ok=True)
        f.write_text(src)
    monkeypatch.syspath_prepend(str(root))
    for m in [k for k in sys.modules if k.split(".")[0] in ("cellpose", "classpose")]:
        monkeypatch.delitem(sys.modules, m)
    import classpose.core as rcore
    import classpose.models as rmodels
    from classpose_b200 import dynamics, models, transforms as btf, utils as butils
    try:
        patched = hooks.install()
        assert {"classpose.models.compute_masks", "classpose.models.compute_class_masks", "classpose.core.unaugment_class_tiles",
                "cellpose.dynamics.resize_and_compute_masks", "cellpose.transforms.average_tiles",
                "cellpose.transforms.unaugment_tiles", "cellpose.utils.fill_holes_and_remove_small_masks"} <= set(patched)
        a, c, b, u = rmodels.ClassposeModel().eval_tail()
        assert a is models.compute_masks and c is models.compute_class_masks
        assert b is dynamics.resize_and_compute_masks and u is butils.fill_holes_and_remove_small_masks
        ua, uc, d = rcore.run_net_tail()
        assert ua is btf.unaugment_tiles and uc is btf.unaugment_class_tiles and d is btf.average_tiles
    finally:
        hooks.uninstall()
    a, c, b, u = rmodels.ClassposeModel().eval_tail()
    assert a() == "stock-b" and c(None, None) == "stock-c" and b() == "stock-b" and u() == "stock-u"
    ua, uc, d = rcore.run_net_tail()
    assert ua() == "stock-ua" and uc(None) == "stock-uc" and d() == "stock-d"
    for m in [k for k in sys.modules if k.split(".")[0] in ("cellpose", "classpose")]:
        sys.modules.pop(m, None)


def test_border_removal_keeps_caller_dtype_and_values(monkeypatch):
    """Host side of remove_border_instances (metrics/pq.py:65-92): only int32 ids travel to the device; float masks and ids
    beyond int32 go through ranks, and the caller's array is edited in place in its own dtype.  The device call is replaced
    by the oracle here (no GPU), which leaves exactly the host logic under test."""
    import parity_cases as pc
    from classpose_b200 import metrics as bm
    from oracle import classpose_ref

    class FakeEngine:
        def remove_border_instances(self, ids, lcap, nch=1):
            assert ids.dtype == np.int32 and nch == 1 and ids.max() < lcap
            return torch.from_numpy(classpose_ref.remove_border_instances(ids[0].copy())[None])
    monkeypatch.setattr(bm, "get_engine", lambda device=None: FakeEngine())
    rng = np.random.default_rng(2)
    base = pc.random_label_image(rng, 40, 56, 12)
    for dtype, scale, off in ((np.float64, 1.5, 0.25), (np.int64, 1, 2 ** 33), (np.uint32, 1, 2 ** 31 + 5), (np.int16, 1, 0)):
        inst = np.where(base > 0, base.astype(np.int64) * scale + off, 0).astype(dtype)
        for nch in (0, 3):
            a = inst.copy() if nch == 0 else np.stack([inst] + [rng.integers(1, 9, size=inst.shape).astype(dtype)
                                                                 for _ in range(nch - 1)], axis=-1)
            want = classpose_ref.remove_border_instances(a.copy())
            got = bm.remove_border_instances(a)
            assert got is a and got.dtype == dtype
            np.testing.assert_array_equal(got, want)


def test_reference_signatures_are_honoured():
    import inspect
    from classpose_b200 import dynamics, models
    sig = inspect.signature(models.compute_masks)
    assert list(sig.parameters) == ["dP", "cellprob", "shape", "do_3D", "niter", "cellprob_threshold", "flow_threshold",
                                    "min_size", "max_size_fraction", "stitch_threshold", "device"]
    sig = inspect.signature(dynamics.resize_and_compute_masks)
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["niter"] == 200 and d["cellprob_threshold"] == 0.0 and d["flow_threshold"] == 0.4
    assert d["min_size"] == 15 and d["max_size_fraction"] == 0.4 and d["resize"] is None
    assert list(inspect.signature(models.compute_class_masks).parameters)[:2] == ["masks", "y_class"]


def test_shard_ranges_cover_the_tile_list():
    for n, w in ((244036, 8), (10, 3), (5, 8), (1024, 4)):
        spans = [cdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
