"""Parity cases shared by the simulator tests (CPU, `-m "not gpu"`) and the GPU tests.

Every case takes a backend `be` exposing the C-ABI call layer with numpy in / numpy out
(tests/backends.py) and compares the kernels with the oracle on identical seeded inputs.
Bars (BASELINE.json north_star): integer / label work bit-exact given identical inputs;
follow_flows >= 99.9 % identical truncated end points (float32 Euler integration is compared
with torch's CPU grid_sample, which itself differs from torch's CUDA grid_sample at this level);
fused path F1 >= 0.995 at IoU 0.5, matched-cell class exact, cell-count delta <= 0.1 %.
"""
from __future__ import annotations

import os

import numpy as np

from oracle import classpose_ref, dynamics, metrics, synth, transforms as otf, utils as outils

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_tile_cache = {}


def std_tile(seed, **kw):
    key = (seed, tuple(sorted(kw.items())))
    if key not in _tile_cache:
        t = synth.make_tile(seed, **kw)
        st = {}
        t["masks_oracle"] = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"], return_stages=st)
        t["stages"] = st
        _tile_cache[key] = t
    return _tile_cache[key]


def adv_tile():
    if "adv" not in _tile_cache:
        t = synth.make_adversarial_tile()
        st = {}
        t["masks_oracle"] = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"], return_stages=st)
        t["stages"] = st
        _tile_cache["adv"] = t
    return _tile_cache["adv"]


def pack_pfinal(stages, H, W):
    pf = np.full((H, W), -1, np.int32)
    ys, xs = stages["inds"]
    pf[ys, xs] = (stages["p_final"][0].astype(np.int32) << 16) | stages["p_final"][1].astype(np.int32)
    return pf


def c32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------------------------ (2)
def case_follow_flows(be):
    tiles = [std_tile(0, H=128, W=128, n_grid=5), std_tile(1), adv_tile(), std_tile(2, H=96, W=160, n_grid=6)]
    tot = same = 0
    for t in tiles:
        H, W = t["cellprob"].shape
        pf, pfl = be.follow_flows(f32(t["dP"][None]), f32(t["cellprob"][None]), 200, 0.0, want_float=True)
        st = t["stages"]
        ys, xs = st["inds"]
        assert (pf[0][t["cellprob"] <= 0] == -1).all()
        py, px = pf[0][ys, xs] >> 16, pf[0][ys, xs] & 0xFFFF
        eq = (py == st["p_final"][0]) & (px == st["p_final"][1])
        tot += len(ys)
        same += int(eq.sum())
        # un-truncated positions: pixels orbiting a sink amplify float32 rounding differences, so the
        # bound is loose (a tenth of a pixel); the truncated end points below are what get_masks consumes
        d = np.maximum(np.abs(pfl[0, 0][ys, xs] - st["p_float"][0]), np.abs(pfl[0, 1][ys, xs] - st["p_float"][1]))
        assert np.mean(d < 0.1) > 0.999 and np.median(d) < 1e-3
    assert same / tot >= 0.999, f"identical truncated end points: {same}/{tot}"


def case_follow_flows_few_iters_exact(be):
    """With few steps rounding cannot amplify: end points must match the oracle exactly."""
    t = std_tile(1)
    st = {}
    fg = t["cellprob"] > 0
    p = dynamics.follow_flows(t["dP"] * fg / 5.0, np.nonzero(fg), 3).int().numpy()
    pf, _ = be.follow_flows(f32(t["dP"][None]), f32(t["cellprob"][None]), 3, 0.0)
    ys, xs = np.nonzero(fg)
    eq = ((pf[0][ys, xs] >> 16) == p[0]) & ((pf[0][ys, xs] & 0xFFFF) == p[1])
    assert eq.mean() > 0.9999


def case_follow_flows_merge_is_exact(be):
    """Trajectory merging must not change a single bit: compare with the plain kernel on batches whose list
    chunks span several tiles, and on many tiny tiles (more than 4 tiles per chunk -> merging is bypassed)."""
    tiles = [std_tile(s) for s in (1, 3, 4)]
    dP = f32(np.stack([t["dP"] for t in tiles])); cp = f32(np.stack([t["cellprob"] for t in tiles]))
    rng = np.random.default_rng(5)
    small_dP = f32(rng.normal(0, 2.0, size=(96, 2, 16, 16))); small_cp = f32(rng.normal(-1.0, 1.0, size=(96, 16, 16)))
    # drift tiles: a constant flow towards each border (plus noise) parks every trajectory on the clamp at +-1 and on
    # the i = -0.5 / L - 0.5 sampling positions -- the corner cases of the packed step's FADD.RM floor and
    # FMNMX.XORSIGN clamp, which must equal floorf / fmin(fmax()) of the scalar kernel
    drift = np.zeros((8, 2, 64, 64), np.float32)
    for k, (vy, vx) in enumerate(((5, 0), (-5, 0), (0, 5), (0, -5), (5, 5), (-5, -5), (5, -5), (-5, 5))):
        drift[k, 0] = vy; drift[k, 1] = vx
    drift = f32(drift + rng.normal(0, 0.7, size=drift.shape)); drift_cp = f32(np.ones((8, 64, 64)))
    try:
        # (the last batch is small enough for the 256-entry chunks of switch 6 on the simulator's 4-SM device as well)
        for a, b in ((dP, cp), (small_dP, small_cp), (drift, drift_cp), (drift[:3], drift_cp[:3])):
            be.set_follow_merge(0)
            p0, f0 = be.follow_flows(a, b, 200, 0.0, want_float=True)
            fg = b > 0
            # two merge points per 256-pixel chunk; trajectory pool with 1024- and with 256-entry chunks (switch 6);
            # TMA-staged plain kernel (GPU)
            for mode, small in ((1, -1), (2, 0), (2, 1), (3, -1)):
                be.set_follow_merge(mode)
                be.set_switch(6, small)
                p1, f1 = be.follow_flows(a, b, 200, 0.0, want_float=True)
                np.testing.assert_array_equal(p0, p1)
                np.testing.assert_array_equal(f0[:, 0][fg], f1[:, 0][fg])
                np.testing.assert_array_equal(f0[:, 1][fg], f1[:, 1][fg])
    finally:
        be.set_follow_merge(-1)
        be.set_switch(6, -1)


def case_follow_flows_large_tiles(be):
    """Tiles around the 2^22-pixel limit of the trajectory-pool kernel's FADD-formed tap index: 2000 x 2000 still takes
    the pool, 2100 x 2100 falls back to the two-point merge kernel; both must equal the plain scalar kernel bit for
    bit (sparse foreground keeps the case fast).  A tile beyond 2^24 padded pixels is refused with CPB_E_RANGE."""
    if be.name == "sim":
        return      # millions of simulated fibres: GPU only
    rng = np.random.default_rng(11)
    for L in (2000, 2100):
        dP = f32(rng.normal(0, 2.5, size=(1, 2, L, L)))
        cp = np.full((1, L, L), -1.0, np.float32)
        for _ in range(60):                                   # 60 foreground blobs of 24 x 24 pixels, some on the border
            y, x = rng.integers(0, L - 24, size=2)
            cp[0, y:y + 24, x:x + 24] = 1.0
        cp[0, :24, :24] = 1.0; cp[0, L - 24:, L - 24:] = 1.0
        try:
            be.set_follow_merge(0)
            p0, f0 = be.follow_flows(dP, cp, 200, 0.0, want_float=True)
            be.set_follow_merge(-1)
            p1, f1 = be.follow_flows(dP, cp, 200, 0.0, want_float=True)
        finally:
            be.set_follow_merge(-1)
        fg = cp > 0
        np.testing.assert_array_equal(p0, p1)
        np.testing.assert_array_equal(f0[:, 0][fg], f1[:, 0][fg])
        np.testing.assert_array_equal(f0[:, 1][fg], f1[:, 1][fg])


# ------------------------------------------------------------------------------------ (3)
def case_get_masks_exact(be):
    for t in (std_tile(0, H=128, W=128, n_grid=5), std_tile(1), adv_tile(), std_tile(2, H=96, W=160, n_grid=6),
              std_tile(5, H=128, W=128, n_grid=16, axes=(2.5, 3.5))):
        H, W = t["cellprob"].shape
        m, cnt = be.get_masks(pack_pfinal(t["stages"], H, W)[None], 0.4)
        ref = t["stages"]["masks_get"]
        np.testing.assert_array_equal(m[0], ref)
        assert cnt[0] == ref.max()


def case_get_masks_plateaus_and_ties(be):
    """Hand-made end points: plateau bins (each is its own seed), equal-count seeds whose windows
    overlap (later raster position wins), a seed at the tile corner, an over-sized label."""
    H, W = 64, 80
    rng = np.random.default_rng(7)
    fg = np.ones((H, W), bool)
    ys, xs = np.nonzero(fg)
    targets = [(0, 0, 40), (10, 10, 30), (10, 11, 30), (10, 14, 30), (30, 30, 12), (30, 33, 12), (33, 30, 25),
               (50, 70, 11), (63, 79, 60), (40, 8, 10), (20, 60, 2200)]
    ty = np.concatenate([np.full(n, y) for y, x, n in targets])
    tx = np.concatenate([np.full(n, x) for y, x, n in targets])
    # halo of low counts (3) around some seeds so that regions actually grow and overlap
    hy, hx = [], []
    for y, x, n in targets[1:7]:
        for dy in range(-3, 4):
            for dx in range(-3, 4):
                if (dy or dx) and 0 <= y + dy < H and 0 <= x + dx < W:
                    hy += [y + dy] * 3
                    hx += [x + dx] * 3
    ty = np.concatenate([ty, np.array(hy)])
    tx = np.concatenate([tx, np.array(hx)])
    n = len(ys)
    assert len(ty) <= n
    rest = n - len(ty)
    ty = np.concatenate([ty, rng.integers(0, H, rest)])
    tx = np.concatenate([tx, rng.integers(0, W, rest)])
    perm = rng.permutation(n)
    p_final = np.stack([ty[perm], tx[perm]]).astype(np.int32)
    ref = dynamics.get_masks(p_final, (ys, xs), (H, W), max_size_fraction=0.4)
    pf = np.zeros((H, W), np.int32)
    pf[ys, xs] = (p_final[0] << 16) | p_final[1]
    m, cnt = be.get_masks(pf[None], 0.4)
    np.testing.assert_array_equal(m[0], ref)
    assert cnt[0] == ref.max() and ref.max() >= 5


def case_get_masks_no_seeds(be):
    H, W = 32, 48
    pf = np.full((H, W), -1, np.int32)
    ys, xs = np.mgrid[0:H, 0:W]
    pf[:] = (ys.astype(np.int32) << 16) | xs.astype(np.int32)   # every pixel stays put: counts of 1, no seed
    m, cnt = be.get_masks(pf[None], 0.4)
    assert not m.any() and cnt[0] == 0


# ------------------------------------------------------------------------------------ (4)
def case_masks_to_flows_exact(be):
    for lab in (std_tile(1)["labels"], synth.adversarial_labels(), std_tile(2, H=96, W=160, n_grid=6)["labels"]):
        mu = be.masks_to_flows(c32(lab[None]), int(lab.max()) + 2)
        ref = dynamics.masks_to_flows(lab)
        assert np.abs(mu[0] - ref).max() <= 1e-12


def corrupt_flows(t, every=4, seed=0):
    """Replace the flows of every `every`-th planted cell by noise so that its flow error is large."""
    rng = np.random.default_rng(seed)
    dP = t["dP"].copy()
    for l in range(1, int(t["labels"].max()) + 1, every):
        m = t["labels"] == l
        dP[:, m] = rng.normal(0, 3.0, size=(2, int(m.sum()))).astype(np.float32)
    return dP


def case_remove_bad_flow_masks_exact(be):
    for t in (std_tile(1), std_tile(3), adv_tile()):
        lab = t["labels"].astype(np.int32)
        dP = corrupt_flows(t)
        err_ref, _ = dynamics.flow_error(lab, dP)
        ref = dynamics.remove_bad_flow_masks(lab.copy(), dP, 0.4)
        out, err = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), int(lab.max()) + 2, 0.4, want_err=True)
        n = len(err_ref)
        assert np.abs(err[0, 1:n + 1] - err_ref).max() < 1e-9
        assert (err_ref > 0.4).sum() >= 1
        np.testing.assert_array_equal(out[0], ref)


def case_flow_qc_fused_equals_unfused(be):
    """Flow error taken inside the diffusion warp (isolated labels) vs k_flow_err on global T (every label), and
    the job queue vs the static label map: same removal set, errors equal to 1e-12 (summation order differs)."""
    SW_QUEUE, SW_QC = 1, 2
    for t in (std_tile(1), std_tile(3), adv_tile(), std_tile(5, H=128, W=128, n_grid=16, axes=(2.5, 3.5))):
        lab = t["labels"].astype(np.int32)
        dP = corrupt_flows(t)
        lcap = int(lab.max()) + 2
        res = {}
        try:
            for queue, reg in ((0, 0), (1, 0)):              # static label map / job queue
                for fused in (0, 1):
                    be.set_switch(SW_QUEUE, queue); be.set_switch(SW_QC, fused)
                    res[(queue, reg, fused)] = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), lcap, 0.4, want_err=True)
                    if fused == 0:
                        res[(queue, reg, "mu")] = be.masks_to_flows(c32(lab[None]), lcap)
        finally:
            be.set_switch(SW_QUEUE, -1); be.set_switch(SW_QC, -1)
        out0, err0 = res[(0, 0, 0)]
        n = int(lab.max())
        for k, v in res.items():
            if k[2] == "mu":
                np.testing.assert_array_equal(v, res[(0, 0, "mu")], err_msg=str(k))   # T itself is bit-identical
                continue
            out, err = v
            np.testing.assert_array_equal(out, out0, err_msg=str(k))
            assert np.abs(err[0, 1:n + 1] - err0[0, 1:n + 1]).max() < 1e-12, k


def near_threshold_flows(lab, rng, thr=0.4):
    """Flows whose per-label error sits at thr + delta for a ladder of deltas: dP / 5 is the label's own unit flow
    rotated by phi with 2 - 2 cos(phi) = thr + delta (every unit-vector pixel then contributes exactly that much)."""
    mu = dynamics.masks_to_flows(lab)
    deltas = [0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4, 1e-3, -1e-3, 1e-2, -1e-2, 0.2, -0.2]
    dP = np.zeros_like(mu)
    for l in range(1, int(lab.max()) + 1):
        m = lab == l
        if not m.any():
            continue
        e = thr + deltas[int(rng.integers(len(deltas)))]
        phi = np.arccos(1.0 - e / 2.0) * (1 if rng.random() < 0.5 else -1)
        c, s_ = np.cos(phi), np.sin(phi)
        dP[0][m] = 5.0 * (c * mu[0][m] - s_ * mu[1][m])
        dP[1][m] = 5.0 * (s_ * mu[0][m] + c * mu[1][m])
    return dP.astype(np.float32)


def unpack_screen_err(err):
    """CPB_QC_SCREEN=2: labels decided by the float32 screen report -(err32 << 32 | bound) bit-packed."""
    bits = err.view(np.int64)
    packed = bits < 0
    b = bits & 0x7FFFFFFFFFFFFFFF
    e32 = (b >> 32).astype(np.uint32).view(np.float32).astype(np.float64)
    bnd = (b & 0xFFFFFFFF).astype(np.uint32).view(np.float32).astype(np.float64)
    return packed, e32, bnd


def case_flow_qc_screen_is_decision_exact(be):
    """The float32 screen in front of the float64 flow check: (a) its proven bound holds, |err32 - err64| <= bound for
    every label it decides; (b) the removal set equals the float64 path's on planted cells, corrupted flows, random
    touching labels and a ladder of errors within 1e-7 .. 1e-2 of the threshold; (c) labels it cannot separate from
    the threshold are left to the float64 kernel (some must be, on the ladder)."""
    SW = 4
    rng = np.random.default_rng(123)
    jobs = []
    for seed in (1, 3, 6, 7):
        t = std_tile(seed)
        lab = t["labels"].astype(np.int32)
        jobs.append((lab, corrupt_flows(t, seed=seed)))
        jobs.append((lab, near_threshold_flows(lab, rng)))
    t = std_tile(5, H=128, W=128, n_grid=16, axes=(2.5, 3.5))
    jobs.append((t["labels"].astype(np.int32), near_threshold_flows(t["labels"].astype(np.int32), rng)))
    t = adv_tile()
    jobs.append((t["labels"].astype(np.int32), corrupt_flows(t)))
    for trial in range(4):
        lab = random_label_image(rng, 64, 80, int(rng.integers(8, 30)), gaps=False)
        u, inv = np.unique(lab, return_inverse=True)
        lab = inv.reshape(lab.shape).astype(np.int32)
        if lab.max() > 0:
            jobs.append((lab, (5.0 * dynamics.masks_to_flows(lab) + rng.normal(0, 1.2, size=(2,) + lab.shape)).astype(np.float32)))
    # a label beyond the warp kernels (44 px wide: block kernels, float64 only) in flat contact with small ones: the small
    # neighbours find no float32 T for it and must fall back together with whatever they touch
    yy, xx = np.mgrid[0:96, 0:128]
    seeds = [(48, 48, 22), (48, 78, 9), (20, 40, 9), (78, 60, 8), (30, 100, 7), (44, 98, 7), (80, 20, 6)]
    d = np.stack([(yy - a) ** 2 + (xx - b_) ** 2 for a, b_, _ in seeds])
    near = d.argmin(0)
    lab = np.where(d.min(0) <= np.array([r * r for _, _, r in seeds])[near], near + 1, 0).astype(np.int32)
    jobs.append((lab, (5.0 * dynamics.masks_to_flows(lab) + rng.normal(0, 0.8, size=(2,) + lab.shape)).astype(np.float32)))
    n_screened = n_left = n_total = 0
    try:
        for lab, dP in jobs:
            lcap = int(lab.max()) + 2
            be.set_switch(SW, 0)
            out0, err0 = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), lcap, 0.4, want_err=True)
            be.set_switch(SW, 1)
            out1, _ = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), lcap, 0.4, want_err=False)
            be.set_switch(SW, 2)
            out2, err2 = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), lcap, 0.4, want_err=True)
            np.testing.assert_array_equal(out1, out0)
            np.testing.assert_array_equal(out2, out0)
            present = np.isin(np.arange(lcap), np.unique(lab)) & (np.arange(lcap) > 0)
            packed, e32, bnd = unpack_screen_err(err2[0])
            packed &= present
            assert (np.abs(e32 - err0[0])[packed] <= bnd[packed]).all(), "screen bound violated"
            # whatever the screen did not decide carries the float64 error itself
            rest = present & ~packed
            assert np.abs(err2[0][rest] - err0[0][rest]).max(initial=0.0) < 1e-12
            n_screened += int(packed.sum()); n_left += int(rest.sum()); n_total += int(present.sum())
    finally:
        be.set_switch(SW, -1)
    assert n_screened > 0.5 * n_total, (n_screened, n_total)
    assert n_left > 0


# ------------------------------------------------------------------------------------ (5)
def nested_rings(H=96, W=96):
    yy, xx = np.mgrid[0:H, 0:W]
    r2 = (yy - 48) ** 2 + (xx - 48) ** 2
    lab = np.zeros((H, W), np.int32)
    lab[(r2 <= 44 ** 2) & (r2 > 38 ** 2)] = 3     # outer ring
    lab[(r2 <= 30 ** 2) & (r2 > 24 ** 2)] = 1     # inner ring (lower id, processed first upstream)
    lab[r2 <= 8 ** 2] = 2                          # core cell
    lab[(yy - 10) ** 2 + (xx - 10) ** 2 <= 9] = 4  # outside everything, larger than min_size
    lab[2:4, 80:83] = 5                            # tiny
    return lab


def case_fill_holes_exact(be):
    labs = [synth.adversarial_labels(), nested_rings(), std_tile(1)["stages"]["masks_qc"].astype(np.int32)]
    # non-contiguous labels with small ones: exercises the positional size filter upstream uses
    q = nested_rings().copy()
    q[q == 4] = 9
    q[q == 5] = 6
    q[60:62, 2:5] = 7
    labs.append(q)
    # all foreground, no background pixel at all
    a = np.ones((40, 40), np.int32)
    a[:, 20:] = 2
    a[5:8, 5:8] = 3
    a[30:32, 30:33] = 5
    labs.append(a)
    for lab in labs:
        for min_size in (15, 0):
            ref = outils.fill_holes_and_remove_small_masks(lab.copy(), min_size=min_size)
            out, cnt = be.fill_holes_and_remove_small_masks(c32(lab[None]).copy(), int(lab.max()) + 2, min_size)
            np.testing.assert_array_equal(out[0], ref, err_msg=f"min_size={min_size}")
            if min_size > 0:
                assert cnt[0] == ref.max()


def case_fill_holes_oversized_label(be):
    """A label whose crop exceeds the shared-memory bitmaps of the block kernel (bbox 600 x 600 > 8192 words): a ring
    with a thick wall, another label and background inside its hole, an L-shaped label with a big bbox and no hole.
    The bitmaps then come from the global pool; the result must equal the oracle (it used to be skipped silently)."""
    H = W = 640
    yy, xx = np.mgrid[0:H, 0:W]
    r2 = (yy - 320) ** 2 + (xx - 320) ** 2
    lab = np.zeros((H, W), np.int32)
    lab[(r2 <= 300 ** 2) & (r2 >= 270 ** 2)] = 1            # ring: bbox 601 x 601
    lab[(np.abs(yy - 320) < 20) & (np.abs(xx - 300) < 30)] = 2     # label inside the hole (gets overwritten)
    lab[(yy < 12) & (xx < 610)] = 3                                  # L shape, bbox 610 x 610, nothing enclosed
    lab[(xx < 12) & (yy < 610)] = 3
    lab[600:630, 600:630] = 4
    lab[610:620, 610:620] = 0                                        # small label with a hole (warp kernel)
    ref = outils.fill_holes_and_remove_small_masks(lab.copy(), 15)
    out, cnt = be.fill_holes_and_remove_small_masks(c32(lab[None]).copy(), int(lab.max()) + 2, 15)
    np.testing.assert_array_equal(out[0], ref)
    assert cnt[0] == ref.max()


def fill_from_input_image(lab):
    """The hole-fill semantic of the kernels, restated for the one configuration in which it differs from upstream:
    every label's holes are taken from the INPUT image (a pixel enclosed by several labels goes to the one with the
    largest bounding box, the outermost), whereas upstream fills label by label in id order on the image as the
    earlier fills left it.  New ids as upstream's loop counter with min_size <= 0: rank of the id among those present."""
    from scipy.ndimage import binary_fill_holes, find_objects
    out = lab.copy()
    best = np.zeros(lab.shape, np.int64)
    for i, slc in enumerate(find_objects(lab)):
        if slc is None:
            continue
        m = lab == (i + 1)
        hole = binary_fill_holes(m) & ~m
        area = (slc[0].stop - slc[0].start) * (slc[1].stop - slc[1].start)
        key = (area << 32) | (i + 1)
        take = hole & (key > best)
        best[take] = key
        out[take] = i + 1
    ids = np.unique(lab); ids = ids[ids != 0]
    lut = np.zeros(int(lab.max()) + 1, lab.dtype)
    lut[ids] = np.arange(1, len(ids) + 1)
    return lut[out]


def case_fill_holes_label_partly_inside_a_hole(be):
    """KNOWN DEVIATION, found by tests/studies/fuzz_sim.py and documented in INTEGRATION.md section 4: label 3 is a ring
    of which one side lies inside a hole of label 2 (which in turn has a piece inside the ring).  Upstream processes
    label 2 first, its fill cuts the ring open, and the ring then encloses nothing; the kernels take both labels' holes
    from the input image, so the ring still fills its interior.  The case pins the kernels' behaviour (and that the two
    differ on exactly the ring's interior), so that a change of either side is noticed."""
    lab = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                    [0, 0, 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                    [0, 12, 12, 12, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                    [0, 0, 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                    [0, 8, 2, 2, 2, 3, 3, 3, 0, 0, 0, 0, 0],
                    [8, 8, 8, 2, 3, 2, 2, 2, 3, 0, 0, 0, 0],
                    [7, 7, 7, 2, 3, 2, 2, 2, 3, 0, 0, 0, 0],
                    [7, 7, 7, 2, 3, 2, 2, 2, 3, 11, 11, 11, 0],
                    [7, 7, 7, 2, 2, 3, 3, 3, 0, 11, 11, 11, 0],
                    [8, 8, 8, 2, 2, 2, 2, 2, 0, 11, 11, 11, 0],
                    [0, 8, 2, 2, 2, 2, 2, 0, 0, 0, 0, 0, 0]], np.int32)
    upstream = outils.fill_holes_and_remove_small_masks(lab.copy(), -1)
    expected = fill_from_input_image(lab)
    out, _ = be.fill_holes_and_remove_small_masks(c32(lab[None]).copy(), int(lab.max()) + 2, -1)
    np.testing.assert_array_equal(out[0], expected)
    differ = out[0] != upstream
    assert differ.sum() == 9 and differ[6:9, 5:8].all()          # the ring's interior, nothing else
    if be.name == "sim":
        # CPB_FILL_EXACT (switch 8, default off): tiles with such a label are replayed label by label -- upstream's result.
        # Validated on the simulator (thousands of tangled images, tests/studies/fuzz_sim.py); it has not run on hardware
        # yet, which is why it is off by default and why this half of the case is simulator-only.
        try:
            be.set_switch(8, 1)
            exact, _ = be.fill_holes_and_remove_small_masks(c32(lab[None]).copy(), int(lab.max()) + 2, -1)
        finally:
            be.set_switch(8, -1)
        np.testing.assert_array_equal(exact[0], upstream)


def random_label_image(rng, H, W, n, gaps=True):
    """Random discs, rings (holes, some with a cell inside), slabs and specks painted over each other, with
    non-contiguous ids when `gaps`: the shapes the table-driven relabelling has to get right."""
    yy, xx = np.mgrid[0:H, 0:W]
    lab = np.zeros((H, W), np.int32)
    ids = rng.permutation(np.arange(1, (2 * n if gaps else n) + 1))[:n]
    for l in ids:
        cy, cx = rng.integers(0, H), rng.integers(0, W)
        kind = rng.integers(0, 5)
        r2 = (yy - cy) ** 2 + (xx - cx) ** 2
        if kind == 0:
            lab[r2 <= rng.integers(2, 9) ** 2] = l
        elif kind == 1:                                    # ring, sometimes with a core of another id painted later
            ro = rng.integers(5, 14); ri = rng.integers(2, ro - 1)
            lab[(r2 <= ro * ro) & (r2 > ri * ri)] = l
        elif kind == 2:                                    # slab touching whatever is there
            h, w = rng.integers(2, 12, size=2)
            lab[cy:cy + h, cx:cx + w] = l
        elif kind == 3:                                    # speck below min_size
            lab[cy:cy + rng.integers(1, 4), cx:cx + rng.integers(1, 4)] = l
        else:                                              # disc with pin holes
            m = r2 <= rng.integers(4, 10) ** 2
            lab[m] = l
            hy, hx = cy + rng.integers(-2, 3), cx + rng.integers(-2, 3)
            if 0 <= hy < H and 0 <= hx < W:
                lab[hy, hx] = 0
    return lab


def case_random_label_images(be):
    """fill_holes_and_remove_small_masks, class vote and border removal on random label images against the oracle /
    the reference's own functions: bit-exact, including nested holes, overwritten cells and the positional size filter."""
    rng = np.random.default_rng(2024)
    for trial in range(24):
        H, W = (int(v) for v in rng.choice([48, 64, 80, 96], size=2))
        lab = random_label_image(rng, H, W, int(rng.integers(3, 40)), gaps=bool(trial % 2))
        lcap = int(lab.max()) + 2
        for min_size in (15, 0):
            ref = outils.fill_holes_and_remove_small_masks(lab.copy(), min_size=min_size)
            out, _ = be.fill_holes_and_remove_small_masks(c32(lab[None]).copy(), lcap, min_size)
            np.testing.assert_array_equal(out[0], ref, err_msg=f"trial {trial} min_size={min_size}")
        C = int(rng.integers(2, 9))
        logits = f32(rng.integers(-2, 3, size=(C, 1, H, W)))          # small integers: plenty of ties
        ref_cm, _ = classpose_ref.compute_class_masks(lab, logits)
        cc, cm = be.class_vote(c32(lab[None]), f32(logits[:, 0][None]), lcap, want_class_masks=True)
        np.testing.assert_array_equal(cm[0].astype(np.int64), ref_cm, err_msg=f"trial {trial} vote")
        ref_b = classpose_ref.remove_border_instances(lab.copy())
        out_b = be.remove_border_instances(c32(lab[None]).copy(), lcap)
        np.testing.assert_array_equal(out_b[0], ref_b, err_msg=f"trial {trial} border")


def case_random_flow_qc(be):
    """remove_bad_flow_masks on random label images full of touching / nested labels (the labels the diffusion warp
    hands over to k_flow_err) with random flows: flow errors to 1e-9, identical removal set; masks_to_flows to 1e-12."""
    rng = np.random.default_rng(77)
    for trial in range(8):
        H, W = (int(v) for v in rng.choice([48, 64, 80], size=2))
        lab = random_label_image(rng, H, W, int(rng.integers(4, 30)), gaps=False)
        # contiguous ids, as get_masks delivers them
        u, inv = np.unique(lab, return_inverse=True)
        lab = inv.reshape(lab.shape).astype(np.int32)
        if lab.max() == 0:
            continue
        mu = dynamics.masks_to_flows(lab)
        dP = (5.0 * mu + rng.normal(0, 1.5, size=mu.shape)).astype(np.float32)
        err_ref, _ = dynamics.flow_error(lab, dP)
        ref = dynamics.remove_bad_flow_masks(lab.copy(), dP, 0.4)
        lcap = int(lab.max()) + 2
        out, err = be.remove_bad_flow_masks(c32(lab[None]).copy(), f32(dP[None]), lcap, 0.4, want_err=True)
        assert np.abs(err[0, 1:len(err_ref) + 1] - err_ref).max() < 1e-9, trial
        decided = np.abs(err_ref - 0.4) > 1e-9          # a label sitting on the threshold may go either way
        bad_ref = err_ref > 0.4
        bad_out = ~np.isin(np.arange(1, len(err_ref) + 1), np.unique(out[0]))
        present = np.isin(np.arange(1, len(err_ref) + 1), np.unique(lab))
        assert ((bad_ref == bad_out) | ~decided | ~present).all(), trial
        if decided.all():
            np.testing.assert_array_equal(out[0], ref, err_msg=f"trial {trial}")
        got = be.masks_to_flows(c32(lab[None]), lcap)
        assert np.abs(got[0] - mu).max() <= 1e-12, trial


# ------------------------------------------------------------------------------------ (6)
def case_class_vote_reference_vectors(be):
    g = np.load(os.path.join(GOLDEN, "ref_class_vote.npz"))
    for k in range(int(g["ncases"])):
        masks, logits = g[f"masks{k}"], g[f"logits{k}"]
        cc, cm = be.class_vote(c32(masks[None]), f32(logits[:, 0][None]), int(masks.max()) + 2, want_class_masks=True)
        np.testing.assert_array_equal(cm[0].astype(np.int64), g[f"class_masks{k}"])
    t = adv_tile()
    m = t["masks_oracle"]
    ref, _ = classpose_ref.compute_class_masks(m, t["logits"][:, None])
    cc, cm = be.class_vote(c32(m[None]), f32(t["logits"][None]), int(m.max()) + 2, want_class_masks=True)
    np.testing.assert_array_equal(cm[0], ref)
    for l in range(1, int(m.max()) + 1):
        assert cc[0, l] == ref[m == l][0]


# ------------------------------------------------------------------------------------ (7)
def case_border_reference_vectors(be):
    g = np.load(os.path.join(GOLDEN, "ref_border.npz"))
    for k in range(int(g["ncases"])):
        a = g[f"in2d_{k}"]
        out = be.remove_border_instances(c32(a[None]).copy(), int(a.max()) + 2, 1)
        np.testing.assert_array_equal(out[0], g[f"out2d_{k}"])
        a = g[f"in3d_{k}"]
        out = be.remove_border_instances(c32(a[None]).copy(), int(a[..., 0].max()) + 2, a.shape[-1])
        np.testing.assert_array_equal(out[0], g[f"out3d_{k}"])


# ------------------------------------------------------------------------------------ (1)
def case_average_tiles(be):
    from classpose_b200 import transforms as btf
    rng = np.random.default_rng(3)
    for augment, nch in ((False, 3), (True, 3), (True, 5)):
        Ly = Lx = 272
        geo = btf.tile_geometry(Ly, Lx, 256, augment=augment, tile_overlap=0.1)
        nt = geo["ny"] * geo["nx"]
        y = rng.normal(size=(nt, nch, 256, 256)).astype(np.float32)
        # oracle: un-augment (flows negate, logits do not), then average
        y5 = y.reshape(geo["ny"], geo["nx"], nch, 256, 256).copy()
        if augment:
            y5 = otf.unaugment_tiles(y5) if nch == 3 else classpose_ref.unaugment_class_tiles(y5)
        ysub = [[a, a + 256] for a in geo["y0"]]
        xsub = [[a, a + 256] for a in geo["x0"]]
        ref = otf.average_tiles(y5.reshape(nt, nch, 256, 256), ysub, xsub, Ly, Lx)
        ty, tx = btf.taper_1d(256, 256)
        out = be.average_tiles(y[None], geo["y0"], geo["x0"], geo["flip"], nch == 3 and augment, ty, tx, Ly, Lx,
                               (0, 0, 0, 0))
        # default vector path: float32 arithmetic with error-free transformations -- within 1e-6 of numpy's float64
        # accumulate everywhere and bit-identical on (far) more than 99.9 % of the elements
        np.testing.assert_allclose(out[0], ref, rtol=1e-6, atol=1e-6)
        assert np.mean(out[0] == ref) > 0.9999, np.mean(out[0] == ref)
        ulp = np.spacing(np.abs(ref).astype(np.float32))
        assert (np.abs(out[0] - ref) <= 4 * ulp).all()      # one ulp of the ACCUMULATOR (up to ~9 x the result), rarely
        crop = (8, 8, 8, 8)
        outc = be.average_tiles(y[None], geo["y0"], geo["x0"], geo["flip"], nch == 3 and augment, ty, tx, Ly, Lx, crop)
        np.testing.assert_array_equal(outc[0], out[0][:, 8:-8, 8:-8])
        # CPB_BLEND_EFT=0: numpy's literal float64 sequence per element; the scalar kernel (no geometry promises) and
        # the 128-bit one are then bit-identical to each other and > 99.9 % identical to the oracle
        try:
            be.set_switch(5, 0)
            out64 = be.average_tiles(y[None], geo["y0"], geo["x0"], geo["flip"], nch == 3 and augment, ty, tx, Ly, Lx, crop)
        finally:
            be.set_switch(5, -1)
        outs = be.average_tiles(y[None], geo["y0"], geo["x0"], geo["flip"], nch == 3 and augment, ty, tx, Ly, Lx, crop,
                                vector=False)
        np.testing.assert_array_equal(outs, out64)
        np.testing.assert_allclose(out64[0], ref[:, 8:-8, 8:-8], rtol=0, atol=1e-6)
        assert np.mean(out64[0] == ref[:, 8:-8, 8:-8]) > 0.999
    # odd geometry: origins not multiples of 4 -> scalar path only
    y = rng.normal(size=(2, 3, 2, 30, 50)).astype(np.float32)
    y0, x0 = np.array([0, 7, 11]), np.array([0, 9, 3])
    ty, tx = np.linspace(0.2, 1, 30), np.linspace(0.3, 1, 50)
    ref = np.zeros((2, 2, 41, 59), np.float32)
    for b in range(2):
        Navg = np.zeros((41, 59)); yf = np.zeros((2, 41, 59), np.float32); m = np.outer(ty, tx)
        for j in range(3):
            yf[:, y0[j]:y0[j] + 30, x0[j]:x0[j] + 50] += y[b, j] * m
            Navg[y0[j]:y0[j] + 30, x0[j]:x0[j] + 50] += m
        with np.errstate(invalid="ignore", divide="ignore"):
            yf /= Navg
        ref[b] = yf
    out = be.average_tiles(y, y0, x0, np.zeros(3, np.int32), False, ty, tx, 41, 59, (0, 0, 0, 0))
    cov = np.isfinite(ref)
    np.testing.assert_allclose(out[cov], ref[cov], rtol=0, atol=1e-6)


# ------------------------------------------------------------------------------------ fused
def fused_compare(be, tiles, C, **kw):
    B = len(tiles)
    dP = f32(np.stack([t["dP"] for t in tiles]))
    cp = f32(np.stack([t["cellprob"] for t in tiles]))
    lg = f32(np.stack([t["logits"] for t in tiles]))
    masks, counts, cell_class, class_masks = be.compute_masks(dP, cp, lg, want_class_masks=True, **kw)
    # without the class image the vote rides on the final label pass (k_final_vote_v4): identical results
    m2, c2, cc2, _ = be.compute_masks(dP, cp, lg, want_class_masks=False, **kw)
    np.testing.assert_array_equal(m2, masks)
    np.testing.assert_array_equal(c2, counts)
    for b in range(B):
        np.testing.assert_array_equal(cc2[b, :int(counts[b]) + 1], cell_class[b, :int(counts[b]) + 1])
    tp = fp = fn = nref = nnew = 0
    for b, t in enumerate(tiles):
        ref = t["masks_oracle"] if not kw else dynamics.resize_and_compute_masks(t["dP"], t["cellprob"], **kw)
        ref_cm, _ = classpose_ref.compute_class_masks(ref, t["logits"][:, None])
        r = metrics.class_agreement(ref, ref_cm, masks[b], class_masks[b].astype(np.int64))
        assert not r["class_mismatch"], f"class differs on matched cells {r['class_mismatch'][:5]}"
        tp += r["tp"]; fp += r["fp"]; fn += r["fn"]; nref += r["n_true"]; nnew += r["n_pred"]
        assert counts[b] == masks[b].max() == len(np.unique(masks[b])) - 1
        for l in range(1, int(counts[b]) + 1):
            assert cell_class[b, l] == class_masks[b][masks[b] == l][0]
    f1 = 1.0 if (2 * tp + fp + fn) == 0 else 2 * tp / (2 * tp + fp + fn)
    return f1, nref, nnew


def case_fused_path(be):
    tiles = [std_tile(s) for s in (1, 3, 4, 6)] + [adv_tile()]
    f1, nref, nnew = fused_compare(be, tiles, 7)
    assert f1 >= 0.995, f1
    assert abs(nnew - nref) <= max(1, 0.001 * nref), (nref, nnew)


def case_fused_path_other_shapes(be):
    tiles = [std_tile(2, H=96, W=160, n_grid=6, C=5), std_tile(7, H=96, W=160, n_grid=6, C=5)]
    f1, nref, nnew = fused_compare(be, tiles, 5)
    assert f1 >= 0.995 and nnew == nref
    dense = [std_tile(5, H=128, W=128, n_grid=16, axes=(2.5, 3.5), C=10)]
    f1, nref, nnew = fused_compare(be, dense, 10)
    assert f1 >= 0.99 and abs(nnew - nref) <= 1


def case_fused_baseline_config_shapes(be):
    """The exact tile shape / class count pairs of BASELINE.json's model configs on 256 x 256 tiles:
    puma (C = 10, configs[2]) and monusac (C = 5, configs[3]); conic (C = 7) is case_fused_path."""
    for C, seeds in ((10, (21, 22)), (5, (23, 24))):
        tiles = [std_tile(s, C=C) for s in seeds]
        f1, nref, nnew = fused_compare(be, tiles, C)
        assert f1 >= 0.995, (C, f1)
        assert abs(nnew - nref) <= max(1, 0.001 * nref), (C, nref, nnew)


def case_eval_tail_blend_fused_with_threshold(be):
    """north_star (1): the tail of ClassposeModel.eval on the network's sub-tile outputs in ONE library call -- the blend of the
    flow map emits the foreground list / scaled flow field / zeroed labels itself (no separate first pass).  Must equal the
    composition `blend -> compute_masks` of the same library bit for bit, and the reference's host sequence on the oracle."""
    from classpose_b200 import transforms as btf
    t = std_tile(23, C=5)
    C = 5
    for augment in (True, False):
        pads = btf.get_pad_yx(256, 256, min_size=(256, 256))
        Ly, Lx = 256 + pads[0] + pads[1], 256 + pads[2] + pads[3]
        geo = btf.tile_geometry(Ly, Lx, 256, augment=augment)
        full = np.zeros((C + 3, Ly, Lx), np.float32)
        full[:C, pads[0]:pads[0] + 256, pads[2]:pads[2] + 256] = t["logits"]
        full[C:C + 2, pads[0]:pads[0] + 256, pads[2]:pads[2] + 256] = t["dP"]
        full[C + 2, pads[0]:pads[0] + 256, pads[2]:pads[2] + 256] = t["cellprob"]
        nt = len(geo["y0"])
        tiles = np.zeros((nt, C + 3, 256, 256), np.float32)
        for j, (y0, x0, f) in enumerate(zip(geo["y0"], geo["x0"], geo["flip"])):
            s_ = full[:, y0:y0 + 256, x0:x0 + 256].copy()
            if f & 1:
                s_ = s_[:, ::-1]; s_[C] *= -1
            if f & 2:
                s_ = s_[:, :, ::-1]; s_[C + 1] *= -1
            tiles[j] = s_
        yfl, ycl = f32(tiles[None, :, C:]), f32(tiles[None, :, :C])
        ty, tx = btf.taper_1d(256, 256)
        masks, counts, cc, cm, dP, cp, lg = be.eval_tail(yfl, ycl, geo["y0"], geo["x0"], geo["flip"], augment, ty, tx, Ly, Lx,
                                                         tuple(int(p) for p in pads), want_class_masks=True)
        # the same library, composed from its parts
        yf = be.average_tiles(yfl, geo["y0"], geo["x0"], geo["flip"], augment, ty, tx, Ly, Lx, tuple(int(p) for p in pads))
        yc = be.average_tiles(ycl, geo["y0"], geo["x0"], geo["flip"], False, ty, tx, Ly, Lx, tuple(int(p) for p in pads))
        np.testing.assert_array_equal(dP[0], yf[0, :2]); np.testing.assert_array_equal(cp[0], yf[0, 2])
        np.testing.assert_array_equal(lg, yc)
        m2, c2, cc2, cm2 = be.compute_masks(f32(yf[:, :2]), f32(yf[:, 2]), f32(yc), want_class_masks=True)
        np.testing.assert_array_equal(masks, m2); np.testing.assert_array_equal(counts, c2)
        np.testing.assert_array_equal(cm, cm2)
        # the reference's host sequence on the oracle
        y5 = tiles[:, C:].reshape(geo["ny"], geo["nx"], 3, 256, 256).copy()
        c5 = tiles[:, :C].reshape(geo["ny"], geo["nx"], C, 256, 256).copy()
        if augment:
            y5 = otf.unaugment_tiles(y5); c5 = classpose_ref.unaugment_class_tiles(c5)
        ysub = [[a, a + 256] for a in geo["y0"]]; xsub = [[a, a + 256] for a in geo["x0"]]
        ryf = otf.average_tiles(y5.reshape(-1, 3, 256, 256), ysub, xsub, Ly, Lx)[:, pads[0]:Ly - pads[1], pads[2]:Lx - pads[3]]
        ryc = otf.average_tiles(c5.reshape(-1, C, 256, 256), ysub, xsub, Ly, Lx)[:, pads[0]:Ly - pads[1], pads[2]:Lx - pads[3]]
        np.testing.assert_allclose(dP[0], ryf[:2], rtol=1e-6, atol=1e-6)
        ref = dynamics.resize_and_compute_masks(ryf[:2], ryf[2])
        ref_cm, _ = classpose_ref.compute_class_masks(ref, ryc[:, None])
        r = metrics.class_agreement(ref, ref_cm, masks[0], cm[0].astype(np.int64))
        assert r["f1"] >= 0.995 and not r["class_mismatch"] and r["n_pred"] == r["n_true"], r


def case_fused_equals_stages_on_odd_tiles(be):
    """Regression (found by tests/studies/fuzz_sim.py): tiles whose pixel count is not a multiple of 4 take the scalar
    final pass, which skipped a filled hole whenever its label kept its number under the final remap.  The fused path
    must equal the composition of the stage entry points bit for bit; tests/golden/fused_odd_tiles.npz holds the four
    inputs that exposed it (27 x 34, 19 x 29, 27 x 23, 31 x 42: noisy flows, ragged foreground)."""
    g = np.load(os.path.join(GOLDEN, "fused_odd_tiles.npz"))
    for k in range(int(g["ncases"])):
        dP, cp = f32(g[f"dP{k}"]), f32(g[f"cp{k}"])
        niter, thr, min_size = int(g[f"kw{k}"][0]), float(g[f"kw{k}"][1]), int(g[f"kw{k}"][2])
        m0, _, _, _ = be.compute_masks(dP[None], cp[None], None, niter=niter, cellprob_threshold=0.0, flow_threshold=thr,
                                       min_size=min_size, max_size_fraction=0.4)
        pf, _ = be.follow_flows(dP[None], cp[None], niter, 0.0)
        m, _ = be.get_masks(pf, 0.4)
        if thr > 0 and m.max() > 0:
            m, _ = be.remove_bad_flow_masks(c32(m).copy(), dP[None], int(m.max()) + 2, thr)
        m, _ = be.fill_holes_and_remove_small_masks(c32(m).copy(), int(m.max()) + 2, min_size)
        np.testing.assert_array_equal(m0, m, err_msg=f"case {k}")
        # and the oracle on the same end points
        ys, xs = np.nonzero(cp > 0)
        pfin = np.stack([pf[0][ys, xs] >> 16, pf[0][ys, xs] & 0xffff]).astype(np.int32)
        ref = dynamics.get_masks(pfin, (ys, xs), cp.shape, max_size_fraction=0.4)
        if thr > 0 and ref.max() > 0:
            ref = dynamics.remove_bad_flow_masks(ref, dP, threshold=thr)
        ref = outils.fill_holes_and_remove_small_masks(ref, min_size)
        np.testing.assert_array_equal(m0[0], ref, err_msg=f"case {k} vs oracle")


def case_fused_generic_class_count(be):
    """A class count that has no specialised final+vote instance (C = 4; the reference's configs use 5, 7, 10) on a
    tile whose pixel count is a multiple of 4, so the vote still rides on the final pass (generic kernel)."""
    tiles = [std_tile(11, H=64, W=64, n_grid=3, C=4), std_tile(12, H=64, W=64, n_grid=3, C=4)]
    f1, nref, nnew = fused_compare(be, tiles, 4)
    assert f1 >= 0.99 and nnew == nref


def case_fused_switch_matrix(be):
    """Every combination of the A/B switches of the fused path (job queue, fused flow error, fused vote, follow_flows
    variant) must give the same label images, counts and cell classes -- they only change how the work is scheduled."""
    tiles = [std_tile(0, H=128, W=128, n_grid=5), adv_tile()]
    outs = {}
    try:
        for t_i, t in enumerate(tiles):
            dP, cp, lg = f32(t["dP"][None]), f32(t["cellprob"][None]), f32(t["logits"][None])
            for follow in (0, 2):
                for queue in (0, 1):
                    for qc in (0, 1):
                        for vote in (0, 1):
                            be.set_follow_merge(follow)
                            be.set_switch(1, queue); be.set_switch(2, qc); be.set_switch(3, vote)
                            # switch 7: seed candidates listed by the counting kernel / found by streaming the histogram
                            be.set_switch(7, (queue + qc + vote) & 1)
                            m, c, cc, _ = be.compute_masks(dP, cp, lg, want_class_masks=False)
                            outs[(t_i, follow, queue, qc, vote)] = (m.copy(), c.copy(), cc[0, :int(c[0]) + 1].copy())
            for follow in (1, 2, 3):          # every counting kernel with the candidate list on and off
                for cands in (0, 1):
                    be.set_follow_merge(follow); be.set_switch(7, cands)
                    m, c, cc, _ = be.compute_masks(dP, cp, lg, want_class_masks=False)
                    outs[(t_i, follow, 9, 9, cands)] = (m.copy(), c.copy(), cc[0, :int(c[0]) + 1].copy())
    finally:
        be.set_follow_merge(-1)
        for sw in (1, 2, 3, 7):
            be.set_switch(sw, -1)
    for t_i in range(len(tiles)):
        m0, c0, cc0 = outs[(t_i, 0, 0, 0, 0)]
        for k, (m, c, cc) in outs.items():
            if k[0] != t_i:
                continue
            np.testing.assert_array_equal(m, m0, err_msg=str(k))
            np.testing.assert_array_equal(c, c0, err_msg=str(k))
            np.testing.assert_array_equal(cc, cc0, err_msg=str(k))


def case_fused_odd_width(be):
    """W not a multiple of 4 (scalar prep kernel, unaligned rows) and a non-multiple-of-32 tile."""
    tiles = [std_tile(9, H=70, W=57, n_grid=3, C=3), std_tile(10, H=70, W=57, n_grid=3, C=3)]
    for t in tiles:
        pf, _ = be.follow_flows(f32(t["dP"][None]), f32(t["cellprob"][None]), 200, 0.0)
        ys, xs = t["stages"]["inds"]
        eq = ((pf[0][ys, xs] >> 16) == t["stages"]["p_final"][0]) & ((pf[0][ys, xs] & 0xFFFF) == t["stages"]["p_final"][1])
        assert eq.mean() >= 0.99     # a sink sitting on an integer boundary flips truncation for its pixels
    f1, nref, nnew = fused_compare(be, tiles, 3)
    assert f1 >= 0.99 and nnew == nref


def case_fused_empty_and_params(be):
    H = W = 64
    dP = np.zeros((2, 2, H, W), np.float32)
    cp = -np.ones((2, H, W), np.float32)
    t = std_tile(0, H=128, W=128, n_grid=5)
    masks, counts, cc, cm = be.compute_masks(dP, cp, None)
    assert not masks.any() and not counts.any()
    # flow check off / no min size / border removal: compare with the oracle under the same switches
    one = [t]
    for kw in (dict(flow_threshold=0.0), dict(min_size=0), dict(cellprob_threshold=1.5), dict(niter=50)):
        f1, nref, nnew = fused_compare(be, one, 7, **kw)
        assert f1 >= 0.99 and abs(nref - nnew) <= 1, kw
    m, c, _, _ = be.compute_masks(f32(t["dP"][None]), f32(t["cellprob"][None]), None, remove_border=True)
    ref = classpose_ref.remove_border_instances(t["masks_oracle"].astype(np.int32).copy())
    r = metrics.match_instances(ref, m[0])
    assert r["fp"] == 0 and r["fn"] == 0


def case_fused_min_size_zero_keeps_upstream_ids(be):
    """resize_and_compute_masks(min_size=0): upstream fills holes but never renumbers by first appearance, so ids follow
    the label order and a label swallowed by a hole leaves a gap.  The fused path must give the same ids (counts = highest id)
    and the stage call must agree with it."""
    t = adv_tile()
    ref = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"], min_size=0)
    assert len(np.unique(ref)) - 1 < ref.max(), "scenario must contain a swallowed label"
    masks, counts, _, _ = be.compute_masks(f32(t["dP"][None]), f32(t["cellprob"][None]), None, min_size=0)
    np.testing.assert_array_equal(masks[0], ref.astype(np.int32))
    assert counts[0] == ref.max()
    for s in (1, 3):
        t = std_tile(s)
        ref = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"], min_size=-1)
        masks, counts, _, _ = be.compute_masks(f32(t["dP"][None]), f32(t["cellprob"][None]), None, min_size=-1)
        r = metrics.match_instances(ref.astype(np.int32), masks[0])
        assert r["f1"] >= 0.995
        same = (masks[0] > 0) == (ref > 0)
        assert same.mean() > 0.999


def case_fused_qc_then_positional_size_filter(be):
    """Fused path where the flow check removes labels first, so the later size filter (which upstream
    indexes by POSITION in the sorted unique list) removes labels other than the small ones."""
    t = dict(std_tile(8))
    t["dP"] = corrupt_flows(t, every=5, seed=3)
    kw = dict(min_size=150)
    st = {}
    ref = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"], return_stages=st, **kw)
    qc = st["masks_qc"]
    assert len(np.unique(qc)) - 1 < qc.max(), "flow check must have removed labels"
    # what a by-label ("intended") size filter would have produced differs from upstream's by-position one
    ids, cnt = np.unique(qc[qc > 0], return_counts=True)
    intended = np.isin(qc, ids[cnt >= 150])
    assert (intended != (ref > 0)).any(), "scenario does not exercise the positional quirk"
    masks, counts, _, _ = be.compute_masks(f32(t["dP"][None]), f32(t["cellprob"][None]), None, **kw)
    r = metrics.match_instances(ref, masks[0])
    assert r["fp"] == 0 and r["fn"] == 0 and r["f1"] == 1.0, {k: v for k, v in r.items() if k != "pairs"}
    assert counts[0] == ref.max()
    assert ((masks[0] > 0) == (ref > 0)).mean() > 0.9995


def contour_test_labels():
    """Label images with blobs, thin lines, necks, diagonal chains, multi-component labels, border cells."""
    rng = np.random.default_rng(11)
    H, W = 96, 128
    yy, xx = np.mgrid[0:H, 0:W]
    lab = np.zeros((H, W), np.int32)
    k = 0
    for _ in range(40):                       # random ellipses (overlaps overwrite -> odd shapes, split labels)
        cy, cx, a, b, th = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(1, 9), rng.uniform(1, 9), rng.uniform(0, 3.2)
        u = ((xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)) / a
        v = (-(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)) / b
        k += 1
        lab[u * u + v * v <= 1] = k
    k += 1; lab[5, 10:30] = k                 # horizontal 1-px line
    k += 1; lab[10:30, 3] = k                 # vertical line
    k += 1
    for i in range(12): lab[40 + i, 60 + i] = k   # diagonal chain
    k += 1; lab[70:74, 20:24] = k; lab[74, 23] = k; lab[75:79, 23:27] = k   # two squares joined by a neck
    k += 1; lab[60:63, 100:103] = k; lab[66:69, 106:109] = k; lab[80, 90] = k   # three components, one label
    k += 1; lab[0:3, 50:55] = k               # touches the top border
    k += 1; lab[H - 1, W - 4:W] = k           # bottom-right corner line
    k += 1; lab[30, 100] = k                  # single pixel
    k += 1; lab[20:25, 110:115] = k; lab[22, 112] = 0   # ring with a hole (outer border only)
    noise = rng.uniform(size=(H, W)) < 0.03   # salt noise creates ragged borders
    lab[noise] = 0
    return lab


def case_cell_contours_match_cv2(be):
    """Next row N1: contours[0] of cv2.findContours(cell_mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) -- the very call
    the reference's PostProcessor makes (predict_wsi.py:614-618) -- plus bbox / area / polygon measures."""
    import cv2
    from scipy.ndimage import find_objects
    labs = [contour_test_labels(), adv_tile()["masks_oracle"].astype(np.int32), std_tile(1)["masks_oracle"].astype(np.int32)]
    for lab in labs:
        H, W = lab.shape
        lcap = int(lab.max()) + 2
        out = be.cell_contours(c32(lab[None]), lcap)
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
            # polygon measures of the closed ring (shoelace), as shapely computes them
            x, y = ref[:, 0].astype(np.float64), ref[:, 1].astype(np.float64)
            xn, yn = np.roll(x, -1), np.roll(y, -1)
            cross = x * yn - xn * y
            assert f[5] == int(round(cross.sum()))
            if len(ref) >= 2:
                np.testing.assert_allclose(out["perimeter"][0, l], np.hypot(xn - x, yn - y).sum(), rtol=1e-12)
            if f[5] != 0:
                cx = ((x + xn) * cross).sum() / (3 * cross.sum()); cy = ((y + yn) * cross).sum() / (3 * cross.sum())
                np.testing.assert_allclose([f[6] / (3 * f[5]), f[7] / (3 * f[5])], [cx, cy], rtol=1e-12)
            # validity: brute-force restatement of "ring neither touches nor crosses itself"
            assert bool(out["valid"][0, l]) == ring_is_simple(ref), f"label {l}: {ref.tolist()}"
        assert int(out["total"][0]) == total
    # a points buffer that is too small reports the needed total and writes only what fits
    lab = labs[0]
    small = be.cell_contours(c32(lab[None]), int(lab.max()) + 2, points_cap=1024)
    full = be.cell_contours(c32(lab[None]), int(lab.max()) + 2)
    assert small["total"][0] == full["total"][0]


def ring_is_simple(pts):
    """Brute-force polygon validity on integer points: >= 4 points, non-zero area, no self contact."""
    n = len(pts)
    if n < 4:
        return False
    P = [tuple(int(v) for v in p) for p in pts]
    a2 = sum(P[i][0] * P[(i + 1) % n][1] - P[(i + 1) % n][0] * P[i][1] for i in range(n))
    if a2 == 0:
        return False

    def orient(a, b, c):
        v = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
        return (v > 0) - (v < 0)

    def on(a, b, p):
        return min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= p[1] <= max(a[1], b[1])

    def touch(a, b, c, d):
        o1, o2, o3, o4 = orient(a, b, c), orient(a, b, d), orient(c, d, a), orient(c, d, b)
        if o1 != o2 and o3 != o4:
            return True
        return (o1 == 0 and on(a, b, c)) or (o2 == 0 and on(a, b, d)) or (o3 == 0 and on(c, d, a)) or (o4 == 0 and on(c, d, b))

    for i in range(n):
        a, b, c = P[i], P[(i + 1) % n], P[(i + 2) % n]
        if orient(a, b, c) == 0 and (b[0] - a[0]) * (c[0] - b[0]) + (b[1] - a[1]) * (c[1] - b[1]) < 0:
            return False
        for j in range(i + 2, n):
            if i == 0 and j == n - 1:
                continue
            if touch(a, b, P[j], P[(j + 1) % n]):
                return False
    return True


def overlap_duplicates(seed=0, n_cells=3000, extent=4000.0):
    """Cells of a slide seen by overlapping tiles: every cell appears 1-4 times with sub-pixel centroid jitter and
    slightly different areas (what the 64-px tile overlap of predict_wsi produces); cells are >= 16 px apart."""
    rng = np.random.default_rng(seed)
    g = int(extent // 16)
    pick = rng.choice(g * g, size=n_cells, replace=False)
    base = np.stack([(pick % g) * 16.0 + rng.uniform(2, 6, n_cells), (pick // g) * 16.0 + rng.uniform(2, 6, n_cells)], 1)
    copies = rng.choice([1, 1, 1, 2, 2, 4], size=n_cells)
    centers = np.concatenate([base[i] + rng.normal(0, 0.4, size=(c, 2)) for i, c in enumerate(copies)])
    sizes = np.concatenate([rng.uniform(80, 300) + rng.uniform(-5, 5, size=c) for c in copies])
    perm = rng.permutation(len(sizes))
    return centers[perm], sizes[perm]


def case_dedup_overlapping_tiles(be):
    """Next row N2: duplicates from tile overlaps (pairs / cliques): identical to the reference's greedy grouping."""
    from oracle import dedup as odedup
    for seed in (0, 1):
        centers, sizes = overlap_duplicates(seed)
        ref = odedup.reference_greedy(centers, sizes, 7.5)
        keep, group = be.dedup_cells(centers[:, 0], centers[:, 1], sizes, 7.5, want_group=True)
        np.testing.assert_array_equal(keep.astype(bool), ref)
        assert ref.sum() == 3000 and len(np.unique(group)) == 3000
    # the reference's order dependence only shows on chains: any visiting order of the pairs gives the same answer here
    pairs = odedup.query_pairs(centers, 7.5)
    np.testing.assert_array_equal(odedup.reference_greedy(centers, sizes, 7.5, pairs[::-1]), ref)


def case_dedup_random_points_components(be):
    """Dense random points (chains and clusters): the device result is the connected-components rule exactly."""
    from oracle import dedup as odedup
    rng = np.random.default_rng(4)
    for n, extent in ((2000, 300.0), (5000, 2000.0), (1, 10.0), (2, 1.0)):
        centers = rng.uniform(0, extent, size=(n, 2)) - extent / 3      # negative coordinates too
        sizes = rng.integers(10, 60, size=n).astype(np.float64)          # many ties -> lowest index must win
        keep, _ = be.dedup_cells(centers[:, 0], centers[:, 1], sizes, 7.5)
        np.testing.assert_array_equal(keep.astype(bool), odedup.components_keep_largest(centers, sizes, 7.5))


def case_prepare_tiles(be):
    """Next row N4: percentile normalisation + pad + make_tiles (with TTA flips) against numpy's own percentile."""
    from classpose_b200 import transforms as btf
    rng = np.random.default_rng(8)
    imgs = []
    a = rng.integers(0, 256, size=(256, 256, 3)).astype(np.float32)            # uint8-valued RGB tile
    a[..., 2] = 137.0                                                          # constant channel: left untouched
    imgs.append(a)
    b = rng.normal(120, 40, size=(225, 225, 3)).astype(np.float32)             # 225-px WSI read (puma, 0.22 mpp)
    b[..., 1] = 5.0 + 1e-4 * rng.uniform(size=(225, 225))                      # p99 - p1 <= 1e-3: zeroed
    imgs.append(b)
    imgs.append((rng.gamma(2.0, 30.0, size=(300, 280, 2))).astype(np.float32))  # larger than bsize: 2x2 tiles
    for img in imgs:
        for augment in (False, True):
            ref, ysub, xsub, pads = otf.prepare_tiles(img, 256, augment=augment)
            H, W, C = img.shape
            geo = btf.tile_geometry(H + pads[0] + pads[1], W + pads[2] + pads[3], 256, augment=augment)
            tiles, lowhigh, code = be.prepare_tiles(f32(img[None]), pads, geo["y0"], geo["x0"], geo["flip"], geo["ly"], geo["lx"])
            assert tiles.shape[1:] == ref.shape
            for c in range(C):
                ch = img[..., c]
                if np.ptp(ch) > 0:
                    lo, hi = np.percentile(ch, 1), np.percentile(ch, 99)
                    assert lowhigh[0, c, 0] == np.float32(lo), (lowhigh[0, c, 0], lo)
                    assert lowhigh[0, c, 1] == np.float32(hi) - np.float32(lo)
                    assert code[0, c] == (1 if np.float32(hi) - np.float32(lo) > 1e-3 else 2)
                else:
                    assert code[0, c] == 0
            np.testing.assert_array_equal(tiles[0], ref)


def case_label_offsets(be):
    counts = np.array([3, 0, 7, 1, 250, 12] * 100, np.int32)
    offs, total = be.label_offsets(counts, 1000)
    ref = 1000 + np.cumsum(counts.astype(np.int64)) - counts
    np.testing.assert_array_equal(offs, ref)
    assert total[0] == counts.sum()


ALL_CASES = [case_follow_flows, case_follow_flows_few_iters_exact, case_follow_flows_merge_is_exact,
             case_follow_flows_large_tiles, case_get_masks_exact,
             case_get_masks_plateaus_and_ties, case_get_masks_no_seeds, case_masks_to_flows_exact,
             case_remove_bad_flow_masks_exact, case_flow_qc_fused_equals_unfused, case_flow_qc_screen_is_decision_exact,
             case_fill_holes_exact, case_fill_holes_oversized_label, case_fill_holes_label_partly_inside_a_hole, case_random_label_images, case_random_flow_qc, case_class_vote_reference_vectors,
             case_border_reference_vectors, case_average_tiles, case_fused_path, case_fused_path_other_shapes, case_fused_baseline_config_shapes, case_fused_equals_stages_on_odd_tiles,
             case_eval_tail_blend_fused_with_threshold,
             case_fused_generic_class_count, case_fused_switch_matrix, case_fused_odd_width, case_fused_empty_and_params, case_fused_min_size_zero_keeps_upstream_ids,
             case_fused_qc_then_positional_size_filter,
             case_cell_contours_match_cv2, case_prepare_tiles, case_dedup_overlapping_tiles, case_dedup_random_points_components,
             case_label_offsets]
