"""Fuzz campaign: the kernel sources on the CPU grid simulator against the oracle, on random small inputs.

    python tests/studies/fuzz_sim.py [seconds] [seed]

Not part of the test suite (it runs for as long as it is given); a bounded slice of it is (tests/test_sim_fuzz.py).
Targets, all exact comparisons:
  get_masks            random end points (clusters, plateaus, borders, odd tile shapes)        vs oracle.dynamics.get_masks
  fill holes + sizes   random label images (nested labels, gaps in the ids, min_size variants)  vs oracle.utils
  class vote           random labels / logits with exact ties                                   vs oracle.classpose_ref
  border removal       random label images, 1 and 3 channels                                    vs oracle.classpose_ref
  flow check           random label images + random flows: removal set                          vs oracle.dynamics
  fused path           every A/B switch combination gives one result                             (library vs itself)
  exact replay         CPB_FILL_EXACT=1: hole fill equal to upstream on tangled label images too, stage call and fused path (planted labels)
  follow_flows         1 .. 4 Euler steps on random flows, any tile shape: truncated end points         vs oracle (torch grid_sample)
  batch consistency    a batch of different tiles equals the tiles one by one                    (library vs itself)
  fused == stages      fused path vs follow -> get_masks -> flow check -> fill through the stage calls  (library vs itself)
  contours             random label images: point lists, area, bbox                             vs cv2.findContours
  masks_to_flows       random label images: <= 1e-12                                            vs oracle.dynamics
  eval_tail            blend fused with the threshold + masks + classes in one call             vs blend -> compute_masks (library)
  prepare_tiles        random images / tile sizes / TTA: bit-exact                              vs oracle.transforms
  dedup                random cell centres (ties, chains): keep flags                          vs oracle.dedup
  average_tiles        random tile geometry / channels / TTA flips: <= 1e-6                     vs oracle.transforms
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import classpose_ref, dynamics, utils as outils  # noqa: E402


def c32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def random_shape(rng, vec_bias=0.5, big=0.0):
    H = int(rng.integers(6, 44))
    W = int(rng.integers(6, 44))
    if rng.random() < big:                        # several 1024-entry chunks of the foreground list, several blocks per pass
        H = int(rng.integers(60, 120)); W = int(rng.integers(60, 120))
    if rng.random() < vec_bias:
        W = max(8, W // 4 * 4)
    return H, W


def random_labels(rng, H, W, nlab):
    """Blobs, rings and nested labels with gaps in the ids; later labels overwrite earlier ones."""
    lab = np.zeros((H, W), np.int32)
    yy, xx = np.mgrid[0:H, 0:W]
    ids = rng.permutation(np.arange(1, 2 * nlab + 1))[:nlab]
    for l in ids:
        cy, cx = rng.integers(0, H), rng.integers(0, W)
        ry, rx = rng.uniform(1.0, max(2.0, H / 3)), rng.uniform(1.0, max(2.0, W / 3))
        d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        kind = rng.random()
        if kind < 0.25:
            sel = (d <= 1.0) & (d >= rng.uniform(0.2, 0.6))            # ring: a hole
        elif kind < 0.35:
            sel = (np.abs(yy - cy) <= 1) & (np.abs(xx - cx) <= rx)     # line
        else:
            sel = d <= 1.0
        lab[sel] = l
    return lab


def fuzz_get_masks(be, rng):
    H, W = random_shape(rng)
    fg = rng.random((H, W)) < rng.uniform(0.3, 1.0)
    if fg.sum() < 4:
        fg[:] = True
    ys, xs = np.nonzero(fg)
    n = len(ys)
    ncl = int(rng.integers(1, 9))
    cy = rng.integers(0, H, ncl); cx = rng.integers(0, W, ncl)
    spread = rng.choice([0, 0, 1, 2])
    which = rng.integers(0, ncl, n)
    ty = np.clip(cy[which] + rng.integers(-spread, spread + 1, n), 0, H - 1)
    tx = np.clip(cx[which] + rng.integers(-spread, spread + 1, n), 0, W - 1)
    stay = rng.random(n) < rng.uniform(0.0, 0.5)            # some pixels end where they started (counts of 1)
    ty = np.where(stay, ys, ty); tx = np.where(stay, xs, tx)
    p_final = np.stack([ty, tx]).astype(np.int32)
    msf = float(rng.choice([0.4, 0.4, 0.05, 1.0]))
    ref = dynamics.get_masks(p_final, (ys, xs), (H, W), max_size_fraction=msf)
    pf = np.full((H, W), -1, np.int32)
    pf[ys, xs] = (p_final[0] << 16) | p_final[1]
    m, cnt = be.get_masks(pf[None], msf)
    np.testing.assert_array_equal(m[0], ref)
    assert cnt[0] == ref.max()


def partly_swallowed(lab):
    """True when some label lies partly -- not wholly -- inside the hole of another label.  Upstream fills label by
    label in id order on the image as the earlier fills left it, so such a label is cut before its own turn; the
    kernels take every label's holes from the input image (outermost label wins).  The two agree unless this
    predicate holds; the fuzz campaign skips these inputs and counts them."""
    from scipy.ndimage import binary_fill_holes
    ids = np.unique(lab); ids = ids[ids != 0]
    for a in ids:
        m = lab == a
        hole = binary_fill_holes(m) & ~m
        if not hole.any():
            continue
        inside = lab[hole]
        for b in np.unique(inside[inside != 0]):
            if 0 < int((inside == b).sum()) < int((lab == b).sum()):
                return True
    return False


SKIPPED = {"partly_swallowed": 0}


def fuzz_fill_holes(be, rng):
    H, W = random_shape(rng)
    lab = random_labels(rng, H, W, int(rng.integers(1, 10)))
    min_size = int(rng.choice([15, 15, 1, 3, 40, 0, -1]))
    probe = outils._drop_small(lab.copy(), min_size) if min_size > 0 else lab
    if partly_swallowed(probe):
        SKIPPED["partly_swallowed"] += 1
        return
    ref = outils.fill_holes_and_remove_small_masks(lab.copy(), min_size)
    out, cnt = be.fill_holes_and_remove_small_masks(c32(lab[None]), int(lab.max()) + 2, min_size)
    np.testing.assert_array_equal(out[0], ref)


def fuzz_class_vote(be, rng):
    H, W = random_shape(rng)
    lab = random_labels(rng, H, W, int(rng.integers(1, 8)))
    lab = outils.renumber(lab) if rng.random() < 0.7 else lab
    C = int(rng.integers(2, 11))          # (a single class axis is squeezed away by the reference itself)
    logits = rng.integers(-3, 4, size=(C, H, W)).astype(np.float32)      # small integers: exact ties are common
    if rng.random() < 0.5:
        logits += rng.normal(0, 0.5, size=logits.shape).astype(np.float32)
    ref_cm, _ = classpose_ref.compute_class_masks(lab, logits[:, None])
    cell_class, cm = be.class_vote(c32(lab[None]), f32(logits[None]), int(lab.max()) + 2, want_class_masks=True)
    np.testing.assert_array_equal(cm[0].astype(np.int64), ref_cm)


def fuzz_border(be, rng):
    H, W = random_shape(rng)
    lab = random_labels(rng, H, W, int(rng.integers(1, 10)))
    nch = int(rng.choice([1, 3]))
    if nch == 1:
        a = lab.copy()
        ref = classpose_ref.remove_border_instances(a.copy())
        out = be.remove_border_instances(c32(a[None]), int(lab.max()) + 2, 1)
        np.testing.assert_array_equal(out[0], ref)
    else:
        a = np.stack([lab, rng.integers(1, 9, size=lab.shape), rng.integers(1, 9, size=lab.shape)], -1).astype(np.int32)
        ref = classpose_ref.remove_border_instances(a.copy())
        out = be.remove_border_instances(c32(a[None]), int(lab.max()) + 2, 3)
        np.testing.assert_array_equal(out[0], ref)


def fuzz_flow_qc(be, rng):
    H, W = random_shape(rng)
    lab = outils.renumber(random_labels(rng, H, W, int(rng.integers(1, 8))))
    if lab.max() == 0:
        return
    mu = dynamics.masks_to_flows(lab)
    # flows between the true ones and noise: errors on both sides of the threshold
    mix = rng.uniform(0.0, 1.0, size=(1, H, W)) < rng.uniform(0.0, 0.6)
    dP = np.where(mix, rng.normal(0, 3.0, size=mu.shape), 5.0 * mu).astype(np.float32)
    thr = float(rng.choice([0.4, 0.4, 0.1, 1.0]))
    ref = dynamics.remove_bad_flow_masks(lab.copy(), dP, threshold=thr)
    out, _ = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), int(lab.max()) + 2, thr)   # (in place on the simulator)
    if rng.random() < 0.5:                                   # the per-label errors themselves (float64 path): <= 1e-9
        err_ref, _ = dynamics.flow_error(lab, dP)
        out2, err = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), int(lab.max()) + 2, thr, want_err=True)
        assert np.abs(err[0, 1:len(err_ref) + 1] - err_ref).max() < 1e-9
        np.testing.assert_array_equal(out2, out)
    # the library keeps the ids of the survivors or renumbers like the reference: compare as sets of pixels per survivor
    np.testing.assert_array_equal(out[0] > 0, ref > 0)
    np.testing.assert_array_equal(outils.renumber(out[0].astype(np.int32)), outils.renumber(ref.astype(np.int32)))


def fuzz_fused(be, rng):
    H, W = random_shape(rng, vec_bias=0.7)
    H, W = max(H, 16), max(W, 16)
    lab = outils.renumber(random_labels(rng, H, W, int(rng.integers(1, 7))))
    mu = dynamics.masks_to_flows(lab) if lab.max() > 0 else np.zeros((2, H, W))
    dP = (5.0 * mu + rng.normal(0, rng.uniform(0.1, 1.5), size=mu.shape)).astype(np.float32)
    cp = (np.where(lab > 0, 4.0, -4.0) + rng.normal(0, 1.5, size=lab.shape)).astype(np.float32)
    C = int(rng.choice([2, 4, 5, 7, 10]))
    lg = rng.normal(0, 1.0, size=(C, H, W)).astype(np.float32)
    kw = dict(niter=int(rng.choice([200, 40, 33])), cellprob_threshold=float(rng.choice([0.0, 0.5, -1.0])),
              flow_threshold=float(rng.choice([0.4, 0.4, 0.0, 1.5])), min_size=int(rng.choice([15, 15, 3, -1, 0])),
              max_size_fraction=float(rng.choice([0.4, 0.4, 0.1])), fill_holes=bool(rng.random() < 0.8))
    if not kw["fill_holes"]:
        kw["min_size"] = -1
    outs = []
    try:
        for follow, sw in ((0, (0, 0, 0, 0, 0, 0)), (2, (1, 1, 1, 1, 1, 1)), (1, (1, 0, 1, 0, 0, 1)), (2, (0, 1, 0, 1, 1, 0))):
            be.set_follow_merge(follow)
            for which, v in zip((1, 2, 3, 4, 6, 7), sw):
                be.set_switch(which, v)
            m, c, cc, _ = be.compute_masks(dP[None], cp[None], lg[None], want_class_masks=False, **kw)
            outs.append((m.copy(), c.copy(), cc[0, :max(int(c[0]), 0) + 1].copy()))
    finally:
        be.set_follow_merge(-1)
        for which in (1, 2, 3, 4, 6, 7):
            be.set_switch(which, -1)
    for m, c, cc in outs[1:]:
        np.testing.assert_array_equal(m, outs[0][0])
        np.testing.assert_array_equal(c, outs[0][1])
        np.testing.assert_array_equal(cc, outs[0][2])


def fuzz_contours(be, rng):
    """Point lists of every label against cv2.findContours(cell, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)[0] on the label's
    crop -- the call the reference's PostProcessor makes -- plus pixel area and bbox."""
    import cv2
    from scipy.ndimage import find_objects
    H, W = random_shape(rng)
    lab = random_labels(rng, H, W, int(rng.integers(1, 10)))
    if rng.random() < 0.3:                                   # specks and diagonal chains
        n = int(rng.integers(1, 12))
        lab[rng.integers(0, H, n), rng.integers(0, W, n)] = int(lab.max()) + 1
    out = be.cell_contours(c32(lab[None]), int(lab.max()) + 2)
    total = 0
    for l, slc in enumerate(find_objects(lab), start=1):
        if slc is None:
            assert out["npoints"][0, l] == 0
            continue
        ys, xs = slc
        cell = lab[ys, xs] == l
        cs = cv2.findContours(np.uint8(cell), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
        ref = cs[0][:, 0] + np.array([xs.start, ys.start])
        n, off = int(out["npoints"][0, l]), int(out["offsets"][0, l])
        assert off == total
        total += n
        np.testing.assert_array_equal(out["points"][off:off + n], ref, err_msg=f"label {l}")
        f = out["feat"][0, l]
        assert f[0] == cell.sum() and (f[1], f[2] + 1, f[3], f[4] + 1) == (ys.start, ys.stop, xs.start, xs.stop)
    assert int(out["total"][0]) == total


def fuzz_masks_to_flows(be, rng):
    H, W = random_shape(rng)
    lab = outils.renumber(random_labels(rng, H, W, int(rng.integers(1, 8))))
    if lab.max() == 0:
        return
    mu = be.masks_to_flows(c32(lab[None]), int(lab.max()) + 2)
    ref = dynamics.masks_to_flows(lab)
    assert np.abs(mu[0] - ref).max() <= 1e-12


def fuzz_average_tiles(be, rng):
    """Taper blend of sub-tiles with random geometry (tile sizes 16 .. 48, images up to ~3 tiles a side, with and
    without the TTA flips, 1 .. 6 channels) against the oracle's float64 accumulate: <= 1e-6."""
    from classpose_b200 import transforms as btf
    from oracle import transforms as otf
    bsize = int(rng.choice([16, 20, 24, 32, 48]))
    augment = bool(rng.random() < 0.5)
    Ly = int(rng.integers(bsize, 3 * bsize)); Lx = int(rng.integers(bsize, 3 * bsize))
    if rng.random() < 0.5:
        Lx = Lx // 4 * 4
    geo = btf.tile_geometry(Ly, Lx, bsize, augment=augment, tile_overlap=float(rng.choice([0.1, 0.25, 0.5])))
    Ly, Lx, ly, lx = geo["Ly"], geo["Lx"], geo["ly"], geo["lx"]
    nt = geo["ny"] * geo["nx"]
    nch = int(rng.integers(1, 7))
    negate = augment and nch == 3 and bool(rng.random() < 0.7)           # flow maps: dY / dX change sign with the flip
    y = rng.normal(size=(nt, nch, ly, lx)).astype(np.float32)
    y5 = y.reshape(geo["ny"], geo["nx"], nch, ly, lx).copy()
    if augment:
        y5 = otf.unaugment_tiles(y5) if negate else classpose_ref.unaugment_class_tiles(y5)
    ysub = [[a, a + ly] for a in geo["y0"]]
    xsub = [[a, a + lx] for a in geo["x0"]]
    ref = _average_tiles_ref(y5.reshape(nt, nch, ly, lx), ysub, xsub, Ly, Lx, ly, lx)
    ty, tx = btf.taper_1d(ly, lx)
    out = be.average_tiles(y[None], geo["y0"], geo["x0"], geo["flip"], negate, ty, tx, Ly, Lx, (0, 0, 0, 0),
                           vector=bool(rng.random() < 0.7))
    np.testing.assert_allclose(out[0], ref, rtol=1e-6, atol=1e-6)


def _average_tiles_ref(y, ysub, xsub, Ly, Lx, ly, lx):
    """cellpose.transforms.average_tiles for any tile size (the oracle's version, like upstream, builds its mask for
    the tile size of y)."""
    from oracle import transforms as otf
    return otf.average_tiles(y, ysub, xsub, Ly, Lx)


def fuzz_prepare_tiles(be, rng):
    """Percentile normalisation + pad + sub-tiles (with the TTA flips) on random small images and tile sizes."""
    from classpose_b200 import transforms as btf
    from oracle import transforms as otf
    bsize = int(rng.choice([16, 32, 48]))
    H = int(rng.integers(8, 3 * bsize)); W = int(rng.integers(8, 3 * bsize)); C = int(rng.integers(1, 4))
    kind = rng.random()
    if kind < 0.4:
        img = rng.integers(0, 256, size=(H, W, C)).astype(np.float32)         # uint8-valued: many equal values
    elif kind < 0.8:
        img = rng.normal(100, 30, size=(H, W, C)).astype(np.float32)
    else:
        img = rng.gamma(2.0, 30.0, size=(H, W, C)).astype(np.float32)
    if rng.random() < 0.3:
        img[..., int(rng.integers(0, C))] = float(rng.integers(0, 200))       # constant channel
    augment = bool(rng.random() < 0.5)
    ref, ysub, xsub, pads = otf.prepare_tiles(img, bsize, augment=augment)
    geo = btf.tile_geometry(H + pads[0] + pads[1], W + pads[2] + pads[3], bsize, augment=augment)
    tiles, lowhigh, code = be.prepare_tiles(f32(img[None]), pads, geo["y0"], geo["x0"], geo["flip"], geo["ly"], geo["lx"])
    assert tiles.shape[1:] == ref.shape
    np.testing.assert_array_equal(tiles[0], ref)


def fuzz_dedup(be, rng):
    """Overlap de-duplication: keep flags against the connected-components rule of the oracle."""
    from oracle import dedup as odedup
    n = int(rng.integers(1, 400))
    extent = float(rng.choice([5.0, 40.0, 200.0, 2000.0]))
    centers = rng.uniform(0, extent, size=(n, 2)) - extent / 3
    if rng.random() < 0.3:
        centers = np.round(centers)                                             # exact ties in distance
    sizes = rng.integers(10, 20, size=n).astype(np.float64)
    md = float(rng.choice([7.5, 7.5, 1.0, 30.0]))
    keep, _ = be.dedup_cells(centers[:, 0], centers[:, 1], sizes, md)
    np.testing.assert_array_equal(keep.astype(bool), odedup.components_keep_largest(centers, sizes, md))


def fuzz_eval_tail(be, rng):
    """cpb_eval_tail_device (blend fused with the threshold, then masks and classes, one call) against the composition
    blend -> compute_masks of the same library, bit for bit, on random small geometries."""
    from classpose_b200 import transforms as btf
    W = int(rng.choice([64, 128])); H = int(rng.integers(16, 72))
    bsize = int(rng.choice([32, 48, 64]))
    augment = bool(rng.random() < 0.5)
    pads = (int(rng.integers(0, 9)), int(rng.integers(0, 9)), int(rng.choice([0, 4, 8])), int(rng.choice([0, 4, 8, 12])))
    Ly, Lx = H + pads[0] + pads[1], W + pads[2] + pads[3]
    geo = btf.tile_geometry(Ly, Lx, bsize, augment=augment, tile_overlap=float(rng.choice([0.1, 0.3])))
    if (geo["Ly"], geo["Lx"]) != (Ly, Lx) or geo["lx"] % 4 or (np.asarray(geo["x0"]) % 4).any():
        return
    ly, lx, nt = geo["ly"], geo["lx"], len(geo["y0"])
    C = int(rng.choice([2, 5, 7]))
    lab = outils.renumber(random_labels(rng, Ly, Lx, int(rng.integers(1, 7))))
    mu = dynamics.masks_to_flows(lab) if lab.max() > 0 else np.zeros((2, Ly, Lx))
    full = np.zeros((C + 3, Ly, Lx), np.float32)
    full[:C] = rng.normal(0, 1, size=(C, Ly, Lx))
    full[C:C + 2] = 5.0 * mu + rng.normal(0, 0.7, size=mu.shape)
    full[C + 2] = np.where(lab > 0, 4.0, -4.0) + rng.normal(0, 1.5, size=lab.shape)
    tiles = np.zeros((nt, C + 3, ly, lx), np.float32)
    for j, (y0, x0, f) in enumerate(zip(geo["y0"], geo["x0"], geo["flip"])):
        s_ = full[:, y0:y0 + ly, x0:x0 + lx].copy()
        if f & 1:
            s_ = s_[:, ::-1]; s_[C] *= -1
        if f & 2:
            s_ = s_[:, :, ::-1]; s_[C + 1] *= -1
        tiles[j] = s_ + rng.normal(0, 0.05, size=s_.shape).astype(np.float32)       # sub-tiles disagree slightly
    yfl, ycl = f32(tiles[None, :, C:]), f32(tiles[None, :, :C])
    ty, tx = btf.taper_1d(ly, lx)
    kw = dict(niter=int(rng.choice([200, 50])), flow_threshold=float(rng.choice([0.4, 0.0])), min_size=int(rng.choice([15, 3])))
    masks, counts, cc, cm, dP, cp, lg = be.eval_tail(yfl, ycl, geo["y0"], geo["x0"], geo["flip"], augment, ty, tx, Ly, Lx,
                                                     pads, want_class_masks=True, **kw)
    yf = be.average_tiles(yfl, geo["y0"], geo["x0"], geo["flip"], augment, ty, tx, Ly, Lx, pads)
    yc = be.average_tiles(ycl, geo["y0"], geo["x0"], geo["flip"], False, ty, tx, Ly, Lx, pads)
    np.testing.assert_array_equal(dP[0], yf[0, :2]); np.testing.assert_array_equal(cp[0], yf[0, 2])
    np.testing.assert_array_equal(lg, yc)
    m2, c2, cc2, cm2 = be.compute_masks(f32(yf[:, :2]), f32(yf[:, 2]), f32(yc), want_class_masks=True, **kw)
    np.testing.assert_array_equal(masks, m2); np.testing.assert_array_equal(counts, c2)
    np.testing.assert_array_equal(cm, cm2)


def stages_composed(be, dP, cp, kw):
    """follow_flows -> get_masks -> remove_bad_flow_masks -> fill_holes_and_remove_small_masks through the stage-by-stage
    entry points (each of them compared with the oracle by its own target)."""
    pf, _ = be.follow_flows(dP[None], cp[None], kw["niter"], kw["cellprob_threshold"])
    m, _ = be.get_masks(pf, kw["max_size_fraction"])
    if kw["flow_threshold"] > 0 and m.max() > 0:
        m, _ = be.remove_bad_flow_masks(c32(m), dP[None], int(m.max()) + 2, kw["flow_threshold"])
    m, _ = be.fill_holes_and_remove_small_masks(c32(m), int(m.max()) + 2, kw["min_size"])
    if kw.get("remove_border"):
        m = be.remove_border_instances(c32(m), int(m.max()) + 2, 1)
    return m


def fuzz_fused_equals_stages(be, rng):
    """The fused path (raw labels + table relabelling, one final pixel pass) against the composition of the stage entry
    points, bit for bit, on messy inputs: noisy flows, ragged foreground, odd tile shapes (pixel counts that are not a
    multiple of 4 take the scalar final pass -- this target found that it skipped a filled hole whose label keeps its
    number under the remap)."""
    H, W = random_shape(rng, vec_bias=0.7, big=0.15)
    H, W = max(H, 16), max(W, 16)
    lab = outils.renumber(random_labels(rng, H, W, int(rng.integers(2, 9 if H < 60 else 30))))
    mu = dynamics.masks_to_flows(lab) if lab.max() > 0 else np.zeros((2, H, W))
    dP = (5.0 * mu + rng.normal(0, rng.uniform(0.3, 2.5), size=mu.shape)).astype(np.float32)
    cp = (np.where(lab > 0, 4.0, -4.0) + rng.normal(0, 2.0, size=lab.shape)).astype(np.float32)
    kw = dict(niter=int(rng.choice([200, 60])), cellprob_threshold=0.0, flow_threshold=float(rng.choice([0.0, 3.0, 0.4])),
              min_size=int(rng.choice([15, 3, -1, 0])), max_size_fraction=0.4)
    if rng.random() < 0.3:
        kw["remove_border"] = True
    if rng.random() < 0.3:                          # other thresholds, a short integration (plain Euler kernel), size caps
        kw["cellprob_threshold"] = float(rng.choice([0.5, -1.0, 2.0]))
        kw["niter"] = int(rng.choice([10, 31, 32, 200]))
        kw["max_size_fraction"] = float(rng.choice([0.4, 0.1, 1.0]))
    C = int(rng.choice([0, 2, 5, 7, 10]))
    lg = rng.normal(0, 1.0, size=(1, C, H, W)).astype(np.float32) if C else None
    want_cm = bool(C and rng.random() < 0.5)
    m0, c0, cc0, cm0 = be.compute_masks(dP[None], cp[None], lg, want_class_masks=want_cm, **kw)
    ms = stages_composed(be, dP, cp, kw)
    np.testing.assert_array_equal(m0, ms)
    if C:                   # classes of the cells: the fused vote (riding on the final pass or not) against the stage call
        cc1, cm1 = be.class_vote(c32(ms), lg, int(ms.max()) + 2, want_class_masks=True)
        n = int(ms.max())
        np.testing.assert_array_equal(cc0[0, 1:n + 1], cc1[0, 1:n + 1])
        if want_cm:
            np.testing.assert_array_equal(cm0, cm1)


def planted_labels(rng):
    """dP / cellprob that push a chosen -- tangled -- label image through get_masks: every pixel of label k jumps in ONE
    Euler step (niter = 1, displacement dP / 5 pixels) onto a target pixel of its own, so the histogram has one bin per
    label and the label image that comes out is the planted one."""
    H = int(rng.integers(20, 44)); W = int(rng.integers(20, 44))
    lab = random_labels(rng, H, W, int(rng.integers(2, 9)))
    ids, cnt = np.unique(lab, return_counts=True)
    for i, c in zip(ids, cnt):
        if i != 0 and c < 11:
            lab[lab == i] = 0                       # a seed needs more than 10 end points
    lab = outils.renumber(lab)
    n = int(lab.max())
    slots = [(3 + 6 * j, 3 + 6 * i) for j in range((H - 4) // 6) for i in range((W - 4) // 6)]
    if n == 0 or len(slots) < n:
        return None
    pick = rng.permutation(len(slots))[:n]
    dP = np.zeros((2, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for k in range(1, n + 1):
        ty, tx = slots[pick[k - 1]]
        m = lab == k
        dP[0][m] = 5.0 * (ty - yy[m] + 0.25); dP[1][m] = 5.0 * (tx - xx[m] + 0.25)
    return dP, np.where(lab > 0, 4.0, -4.0).astype(np.float32)


def fuzz_fused_planted_labels(be, rng):
    """Tangled label images through the FUSED path against the oracle on the same end points: exact without the
    replay unless a label lies partly inside another label's hole, always exact with it (switch 8, CPB_FILL_EXACT)."""
    g = planted_labels(rng)
    if g is None:
        return
    dP, cp = g
    H, W = cp.shape
    kw = dict(niter=1, cellprob_threshold=0.0, flow_threshold=0.0, min_size=int(rng.choice([-1, 0, 3, 15])), max_size_fraction=1.0)
    try:
        be.set_switch(8, 0); m0, _, _, _ = be.compute_masks(dP[None], cp[None], None, **kw)
        be.set_switch(8, 1); m1, _, _, _ = be.compute_masks(dP[None], cp[None], None, **kw)
        ms = stages_composed(be, dP, cp, kw)
    finally:
        be.set_switch(8, -1)
    pf, _ = be.follow_flows(dP[None], cp[None], 1, 0.0)
    ys, xs = np.nonzero(cp > 0)
    pfin = np.stack([pf[0][ys, xs] >> 16, pf[0][ys, xs] & 0xffff]).astype(np.int32)
    mo = dynamics.get_masks(pfin, (ys, xs), (H, W), max_size_fraction=1.0)
    tangled = partly_swallowed(outils._drop_small(mo.copy(), kw["min_size"]) if kw["min_size"] > 0 else mo)
    ref = outils.fill_holes_and_remove_small_masks(mo, kw["min_size"])
    np.testing.assert_array_equal(m1[0], ref)
    np.testing.assert_array_equal(m1, ms)
    if tangled:
        SKIPPED["partly_swallowed"] += 1
    else:
        np.testing.assert_array_equal(m0[0], ref)


def fuzz_fill_holes_exact_replay(be, rng):
    """fill_holes_and_remove_small_masks with the sequential replay on (switch 8): equal to upstream on EVERY label
    image, tangled ones included."""
    H, W = random_shape(rng)
    lab = random_labels(rng, H, W, int(rng.integers(1, 10)))
    min_size = int(rng.choice([15, 15, 1, 3, 40, 0, -1]))
    ref = outils.fill_holes_and_remove_small_masks(lab.copy(), min_size)
    try:
        be.set_switch(8, 1)
        out, _ = be.fill_holes_and_remove_small_masks(c32(lab[None]), int(lab.max()) + 2, min_size)
    finally:
        be.set_switch(8, -1)
    np.testing.assert_array_equal(out[0], ref)


def fuzz_batch_consistency(be, rng):
    """A batch of 2 .. 4 different tiles through the fused path (with classes) equals the tiles one by one: label images,
    counts, and the class of every cell."""
    H, W = random_shape(rng, vec_bias=0.7)
    H, W = max(H, 16), max(W, 16)
    B = int(rng.integers(2, 5)); C = int(rng.choice([2, 5, 7]))
    dP = np.zeros((B, 2, H, W), np.float32); cp = np.zeros((B, H, W), np.float32)
    for b in range(B):
        lab = outils.renumber(random_labels(rng, H, W, int(rng.integers(0, 8)))) if rng.random() < 0.9 else np.zeros((H, W), np.int32)
        mu = dynamics.masks_to_flows(lab) if lab.max() > 0 else np.zeros((2, H, W))
        dP[b] = 5.0 * mu + rng.normal(0, rng.uniform(0.2, 1.5), size=mu.shape)
        cp[b] = np.where(lab > 0, 4.0, -4.0) + rng.normal(0, 1.5, size=lab.shape)
    lg = rng.normal(0, 1.0, size=(B, C, H, W)).astype(np.float32)
    kw = dict(niter=int(rng.choice([200, 50])), flow_threshold=float(rng.choice([0.4, 0.0, 2.0])), min_size=int(rng.choice([15, 3, -1])))
    m, c, cc, cm = be.compute_masks(dP, cp, lg, want_class_masks=bool(rng.random() < 0.5), **kw)
    for b in range(B):
        m1, c1, cc1, cm1 = be.compute_masks(dP[b:b + 1], cp[b:b + 1], lg[b:b + 1], want_class_masks=cm is not None, **kw)
        np.testing.assert_array_equal(m[b], m1[0]); assert c[b] == c1[0]
        n = max(int(c1[0]), 0)
        np.testing.assert_array_equal(cc[b, :n + 1], cc1[0, :n + 1])
        if cm is not None:
            np.testing.assert_array_equal(cm[b], cm1[0])


def fuzz_follow_flows_few_steps(be, rng):
    """Euler integration against the oracle (torch-CPU grid_sample) for 1 .. 4 steps, where rounding cannot amplify:
    truncated end points equal (at most one pixel in a thousand may sit on a truncation boundary), any tile shape --
    widths that are not a multiple of 4 take the scalar prep kernel."""
    H, W = random_shape(rng)
    H, W = max(H, 8), max(W, 8)
    dP = rng.normal(0, rng.uniform(0.5, 6.0), size=(2, H, W)).astype(np.float32)
    cp = rng.normal(0.5, 1.5, size=(H, W)).astype(np.float32)
    niter = int(rng.integers(1, 5))
    fg = cp > 0
    if fg.sum() == 0:
        return
    p = dynamics.follow_flows(dP * fg / 5.0, np.nonzero(fg), niter).int().numpy()
    pf, _ = be.follow_flows(f32(dP[None]), f32(cp[None]), niter, 0.0)
    ys, xs = np.nonzero(fg)
    assert (pf[0][~fg] == -1).all()
    eq = ((pf[0][ys, xs] >> 16) == p[0]) & ((pf[0][ys, xs] & 0xFFFF) == p[1])
    assert (~eq).sum() <= max(1, len(ys) // 1000), (int((~eq).sum()), len(ys))


TARGETS = [fuzz_follow_flows_few_steps, fuzz_batch_consistency, fuzz_fill_holes_exact_replay, fuzz_fused_planted_labels, fuzz_fused_equals_stages, fuzz_eval_tail, fuzz_prepare_tiles, fuzz_dedup, fuzz_average_tiles, fuzz_contours, fuzz_masks_to_flows, fuzz_get_masks, fuzz_fill_holes, fuzz_class_vote, fuzz_border, fuzz_flow_qc, fuzz_fused]


def run(seconds=60.0, seed=0, targets=TARGETS, be=None, verbose=True):
    if be is None:
        from backends import SimBackend
        be = SimBackend()
    t0 = time.time()
    counts = {f.__name__: 0 for f in targets}
    it = 0
    while time.time() - t0 < seconds:
        f = targets[it % len(targets)]
        rng = np.random.default_rng([seed, it])
        try:
            f(be, rng)
        except Exception:
            print(f"FAILED: {f.__name__} with rng seed [{seed}, {it}]", flush=True)
            raise
        counts[f.__name__] += 1
        it += 1
    if verbose:
        print("ok:", counts, "skipped:", SKIPPED, flush=True)
    return counts


if __name__ == "__main__":
    run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
