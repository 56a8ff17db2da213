"""GPU tests of the reference-facing python boundary (hooks A-D): numpy in, numpy out, same return
types as the reference, results against the oracle / the reference's golden vectors."""
import os

import numpy as np
import pytest

import parity_cases as pc
from oracle import classpose_ref, dynamics as odyn, metrics, transforms as otf, utils as outils

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded():
    from classpose_b200.engine import get_engine
    eng = get_engine()
    n0 = eng.launch_count()
    eng.compute_masks_batch(np.zeros((1, 2, 32, 32), np.float32), np.ones((1, 32, 32), np.float32))
    assert eng.launch_count() > n0
    loaded = open("/proc/self/maps").read()
    assert "libclasspose_b200.so" in loaded


def test_hook_b_resize_and_compute_masks_numpy_contract():
    from classpose_b200 import dynamics
    t = pc.std_tile(1)
    m = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"], niter=200, cellprob_threshold=0.0,
                                          flow_threshold=0.4, min_size=15, max_size_fraction=0.4, resize=None,
                                          device=None)
    assert isinstance(m, np.ndarray) and m.dtype == np.uint16 and m.shape == t["cellprob"].shape
    r = metrics.match_instances(t["masks_oracle"], m)
    assert r["f1"] >= 0.995
    z = dynamics.resize_and_compute_masks(t["dP"], -np.abs(t["cellprob"]) - 1)
    assert z.shape == m.shape and not z.any()
    m2 = dynamics.compute_masks(t["dP"], t["cellprob"])      # min_size=-1: no fill / size filter
    ref2 = odyn.compute_masks(t["dP"], t["cellprob"])
    assert metrics.match_instances(ref2, m2)["f1"] >= 0.995


def test_hook_a_models_compute_masks():
    import torch
    from classpose_b200 import models
    tiles = [pc.std_tile(1), pc.std_tile(3)]
    dP = np.stack([t["dP"] for t in tiles], 1)           # [2, nimg, H, W]
    cp = np.stack([t["cellprob"] for t in tiles])
    one = models.compute_masks(dP[:, :1], cp[:1], (1, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, torch.device("cpu"))
    assert one.shape == (256, 256)
    assert metrics.match_instances(tiles[0]["masks_oracle"], one)["f1"] >= 0.995
    both = models.compute_masks(dP, cp, (2, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, torch.device("cuda"))
    assert both.shape == (2, 256, 256)
    np.testing.assert_array_equal(both[0], one)
    with pytest.raises(NotImplementedError):
        models.compute_masks(dP, cp, (2, 256, 256), True, 200, 0.0, 0.4, 15, 0.4, 0.0, None)


def test_hook_c_compute_class_masks_reference_vectors(golden_dir):
    from classpose_b200 import models
    g = np.load(os.path.join(golden_dir, "ref_class_vote.npz"))
    for k in range(int(g["ncases"])):
        cm, uniq = models.compute_class_masks(g[f"masks{k}"].copy(), g[f"logits{k}"].copy())
        assert cm.dtype == np.int64
        np.testing.assert_array_equal(cm, g[f"class_masks{k}"])
        np.testing.assert_array_equal(uniq, g[f"unique{k}"])
    cm, uniq = models.compute_class_masks(np.zeros((8, 8), np.uint16), np.zeros((3, 1, 8, 8), np.float32))
    assert not cm.any() and list(uniq) == [0]


def test_hook_d_average_tiles_and_unaugment():
    from classpose_b200 import transforms as btf
    rng = np.random.default_rng(0)
    img = rng.normal(size=(3, 272, 272)).astype(np.float32)
    IMG, ysub, xsub, Ly, Lx = otf.make_tiles(img, bsize=256, augment=False)
    y = IMG.reshape(-1, 3, 256, 256)
    out = btf.average_tiles(y, ysub, xsub, Ly, Lx)
    ref = otf.average_tiles(y, ysub, xsub, Ly, Lx)
    assert out.dtype == np.float32 and out.shape == ref.shape
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-6)
    y5 = rng.normal(size=(3, 3, 3, 64, 64)).astype(np.float32)
    np.testing.assert_array_equal(btf.unaugment_tiles(y5.copy()), otf.unaugment_tiles(y5.copy()))
    np.testing.assert_array_equal(btf.unaugment_class_tiles(y5.copy()), classpose_ref.unaugment_class_tiles(y5.copy()))


def test_border_and_fill_wrappers(golden_dir):
    from classpose_b200 import metrics as bmetrics, utils as butils
    g = np.load(os.path.join(golden_dir, "ref_border.npz"))
    for k in range(int(g["ncases"])):
        for tag in ("2d", "3d"):
            a = g[f"in{tag}_{k}"].copy()
            out = bmetrics.remove_border_instances(a)
            assert out is a
            np.testing.assert_array_equal(out, g[f"out{tag}_{k}"])
    lab = pc.nested_rings().astype(np.uint16)
    np.testing.assert_array_equal(butils.fill_holes_and_remove_small_masks(lab.copy(), 15),
                                  outils.fill_holes_and_remove_small_masks(lab.copy(), 15))
    # dtypes the int32 device path cannot hold: surviving pixels -- every channel -- must keep their exact values
    rng = np.random.default_rng(2)
    base = pc.random_label_image(rng, 40, 56, 12)
    for dtype, scale, off in ((np.float64, 1.5, 0.25), (np.int64, 1, 2 ** 33), (np.uint32, 1, 2 ** 31 + 5), (np.int16, 1, 0)):
        inst = np.where(base > 0, base.astype(np.int64) * scale + off, 0).astype(dtype)
        for nch in (0, 3):
            a = inst.copy() if nch == 0 else np.stack([inst] + [rng.integers(1, 9, size=inst.shape).astype(dtype) for _ in range(nch - 1)], axis=-1)
            want = classpose_ref.remove_border_instances(a.copy())
            got = bmetrics.remove_border_instances(a)
            assert got is a and got.dtype == dtype
            np.testing.assert_array_equal(got, want)


def test_host_buffer_path_matches_device_path():
    import torch
    from classpose_b200.engine import get_engine
    eng = get_engine()
    tiles = [pc.std_tile(s) for s in (1, 3, 4)]
    dP = np.stack([t["dP"] for t in tiles]); cp = np.stack([t["cellprob"] for t in tiles])
    lg = np.stack([t["logits"] for t in tiles])
    md, cd, ccd, _ = eng.compute_masks_batch(dP, cp, lg)
    mh, ch, cch, cmh = eng.compute_masks_host(dP, cp, lg, tiles_per_chunk=2, want_class_masks=True)
    np.testing.assert_array_equal(mh.numpy(), md.cpu().numpy())
    np.testing.assert_array_equal(ch.numpy(), cd.cpu().numpy())
    for b in range(3):
        n = int(ch[b])
        np.testing.assert_array_equal(cch[b, :n + 1].numpy(), ccd[b, :n + 1].cpu().numpy())


def test_host_buffer_path_mapped_logits_and_u16_masks():
    """Pinned logits are read in place by the final label pass (no upload); pageable ones are uploaded; uint16 masks;
    the speculative cell_class row width is widened when a tile holds more cells than it (forced by a tiny width)."""
    import torch
    from classpose_b200 import ClassposeB200Error
    from classpose_b200.engine import get_engine
    eng = get_engine()
    tiles = [pc.std_tile(s) for s in (1, 3, 4, 6, 7)]
    dP = torch.from_numpy(np.stack([t["dP"] for t in tiles])); cp = torch.from_numpy(np.stack([t["cellprob"] for t in tiles]))
    lg = torch.from_numpy(np.stack([t["logits"] for t in tiles]))
    md, cd, ccd, _ = eng.compute_masks_batch(dP, cp, lg)
    md, cd, ccd = md.cpu().numpy(), cd.cpu().numpy(), ccd.cpu().numpy()
    pin = [x.pin_memory() for x in (dP, cp, lg)]
    for mode, args in (("mapped", pin), ("auto", pin), ("upload", pin), ("auto", (dP, cp, lg))):
        for u16 in (False, True):
            mh, ch, cch, _ = eng.compute_masks_host(*args, tiles_per_chunk=2, logits_mode=mode, flows_mode=mode, masks_u16=u16)
            assert mh.dtype == (torch.uint16 if u16 else torch.int32)
            np.testing.assert_array_equal(mh.numpy().astype(np.int32), md, err_msg=f"{mode} u16={u16}")
            np.testing.assert_array_equal(ch.numpy(), cd)
            for b in range(len(tiles)):
                n = int(ch[b])
                np.testing.assert_array_equal(cch[b, :n + 1].numpy(), ccd[b, :n + 1])
    with pytest.raises(ClassposeB200Error):
        eng.compute_masks_host(dP, cp, lg, logits_mode="mapped")          # pageable logits cannot be mapped
    with pytest.raises(ClassposeB200Error):
        eng.compute_masks_host(dP, cp, lg, flows_mode="mapped")


def test_two_devices_in_one_process():
    """Function attributes (dynamic shared memory of the block hole-fill / vote kernels) and the SM count are per
    device: a second engine on a second GPU of the same process must work, host path included."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from classpose_b200.engine import get_engine
    t = pc.std_tile(1)
    lab = pc.nested_rings().astype(np.int32)
    for d in (0, 1):
        eng = get_engine(f"cuda:{d}")
        m, c, cc, cm = eng.compute_masks_batch(t["dP"][None], t["cellprob"][None], t["logits"][None], want_class_masks=True)
        assert metrics.match_instances(t["masks_oracle"], m[0].cpu().numpy())["f1"] >= 0.995
        out, _ = eng.fill_holes_and_remove_small_masks(lab[None], int(lab.max()) + 2, 15)
        np.testing.assert_array_equal(out[0].cpu().numpy(), outils.fill_holes_and_remove_small_masks(lab.copy(), 15))
        mh, ch, _, _ = eng.compute_masks_host(t["dP"][None], t["cellprob"][None], t["logits"][None])
        np.testing.assert_array_equal(mh.numpy(), m.cpu().numpy())


def test_single_tile_fast_path_graph_and_identity_cache():
    """Hooks A / B and C as the WSI loop calls them (one tile, numpy in / out): the CUDA-graph plan must give the same
    label image as the batched engine call, hook C must find the labels of the array hook A returned on the device
    (identity cache) and give the same classes as a fresh vote on a copy of that array; a second tile through the
    same plan must not leak state from the first."""
    from classpose_b200 import dynamics, fastpath, models
    from classpose_b200.engine import get_engine
    eng = get_engine()
    for seeds, C in (((1, 3, 4), 7), ((21,), 10)):
        for s_ in seeds:
            t = pc.std_tile(s_, C=C) if C != 7 else pc.std_tile(s_)
            ref_m, ref_c, _, _ = eng.compute_masks_batch(t["dP"][None], t["cellprob"][None])
            ref_m = ref_m[0].cpu().numpy()
            m = models.compute_masks(t["dP"][:, None], t["cellprob"][None], (1, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, None)
            assert m.dtype == np.uint16 and m.shape == (256, 256)
            np.testing.assert_array_equal(m.astype(np.int32), ref_m)
            assert fastpath.plan_holding(m) is not None
            cm, uniq = models.compute_class_masks(m, t["logits"][:, None])             # cached labels, logits only
            assert fastpath.plan_holding(m.copy()) is None
            cm2, uniq2 = models.compute_class_masks(m.copy(), t["logits"][:, None])    # a copy: the general path
            np.testing.assert_array_equal(cm, cm2)
            np.testing.assert_array_equal(uniq, uniq2)
            assert cm.dtype == np.int64 and uniq.dtype == m.dtype
            ref_cm, ref_u = classpose_ref.compute_class_masks(m, t["logits"][:, None])
            np.testing.assert_array_equal(cm, ref_cm)
            np.testing.assert_array_equal(uniq, ref_u)
    # the array of an older call is not confused with the current device labels
    t1, t2 = pc.std_tile(1), pc.std_tile(3)
    m1 = dynamics.resize_and_compute_masks(t1["dP"], t1["cellprob"])
    m2 = dynamics.resize_and_compute_masks(t2["dP"], t2["cellprob"])
    assert fastpath.plan_holding(m1) is None and fastpath.plan_holding(m2) is not None
    cm1, _ = models.compute_class_masks(m1, t1["logits"][:, None])
    np.testing.assert_array_equal(cm1, classpose_ref.compute_class_masks(m1, t1["logits"][:, None])[0])
    # empty tile
    z = dynamics.resize_and_compute_masks(np.zeros((2, 256, 256), np.float32), -np.ones((256, 256), np.float32))
    assert not z.any()
    cmz, uz = models.compute_class_masks(z, t1["logits"][:, None])
    assert not cmz.any() and uz.tolist() == [0]


def test_tile_plans_from_two_threads():
    """The reference's two inference threads (predict_wsi.py:728-797), each calling hook A then hook C per tile: the
    per-thread plans (CUDA graphs on their own streams) must not disturb each other, also while plans are being created."""
    import threading
    from classpose_b200 import models
    tiles = [pc.std_tile(s_) for s_ in (1, 3, 4, 6)]
    want = []
    for t in tiles:
        m = models.compute_masks(t["dP"][:, None], t["cellprob"][None], (1, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, None)
        cm, _ = models.compute_class_masks(m, t["logits"][:, None])
        want.append((m.copy(), cm.copy()))
    errors = []

    def work(k):
        try:
            for rep in range(6):
                i = (k + rep) % len(tiles)
                t = tiles[i]
                m = models.compute_masks(t["dP"][:, None], t["cellprob"][None], (1, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, None)
                cm, _ = models.compute_class_masks(m, t["logits"][:, None])
                np.testing.assert_array_equal(m, want[i][0])
                np.testing.assert_array_equal(cm, want[i][1])
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    [t_.start() for t_ in th]; [t_.join() for t_ in th]
    assert not errors, errors


def test_touching_workload_screen_equals_float64_path():
    """The hostile generator (Voronoi-clipped touching cells, 15 % of them with noise for flows, a 43-px cell every 8th tile,
    a ring every 16th): the float32 flow-check screen -- isolated labels in registers, labels in contact through the
    float32 T plane -- must give exactly the label images of the float64-only path, and must actually decide most labels."""
    import torch
    from classpose_b200 import synth
    from classpose_b200.engine import get_engine
    eng = get_engine()
    data = synth.make_batch(48, 256, 256, 7, seed=5, device=eng.device, style="touching", chunk=16)
    dP, cp, lg = data["dP"], data["cellprob"], data["logits"]
    try:
        eng.lib.cpb_debug_set_switch(4, 0)
        m0, c0, cc0, _ = eng.compute_masks_batch(dP, cp, lg)
        torch.cuda.synchronize()
        eng.lib.cpb_debug_set_switch(4, 1)
        m1, c1, cc1, _ = eng.compute_masks_batch(dP, cp, lg)
        torch.cuda.synchronize()
    finally:
        eng.lib.cpb_debug_set_switch(4, -1)
    assert torch.equal(m0, m1) and torch.equal(c0, c1) and torch.equal(cc0, cc1)
    _, qc = eng.profile_stages(dP, cp, lg, with_qc=True)
    assert qc["screen_decided"] > 2 * qc["float64_labels"], qc
    assert int(c1.sum()) > 48 * 40           # the tiles do hold cells after the flow check
    # and the oracle agrees on a few of them
    for b in (0, 3, 5):
        ref = odyn.resize_and_compute_masks(dP[b].cpu().numpy(), cp[b].cpu().numpy())
        r = metrics.match_instances(ref, m1[b].cpu().numpy())
        assert r["f1"] >= 0.99, (b, r)


def test_concurrent_calls_from_two_threads():
    """The reference runs two inference threads per process; calls must be re-entrant."""
    import threading
    import torch
    from classpose_b200.engine import get_engine
    eng = get_engine()
    tiles = [pc.std_tile(1), pc.std_tile(3)]
    res = [None, None]

    def work(i):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                m, c, _, _ = eng.compute_masks_batch(tiles[i]["dP"][None], tiles[i]["cellprob"][None])
            s.synchronize()
        res[i] = m[0].cpu().numpy()
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    for i in range(2):
        assert metrics.match_instances(tiles[i]["masks_oracle"], res[i])["f1"] >= 0.995


def test_device_synth_and_batch_consistency():
    """Device-side generator + a batch of 64 tiles: every tile recovers its planted cells."""
    import torch
    from classpose_b200 import synth
    from classpose_b200.engine import get_engine
    eng = get_engine()
    d = synth.make_batch(64, 256, 256, 7, seed=5)
    masks, counts, cc, _ = eng.compute_masks_batch(d["dP"], d["cellprob"], d["logits"])
    torch.cuda.synchronize()
    planted = torch.stack([(d["labels"][b].unique() > 0).sum() for b in range(64)])
    assert (counts.cpu() - planted.cpu()).abs().float().mean() < 1.0
    # oracle on two of them
    for b in (0, 63):
        ref = odyn.resize_and_compute_masks(d["dP"][b].cpu().numpy(), d["cellprob"][b].cpu().numpy())
        assert metrics.match_instances(ref, masks[b].cpu().numpy())["f1"] >= 0.995


def test_batch_parts_with_ragged_sizes():
    """The C API cuts batches of >= 256 tiles into up to four parts that run on forked streams (run_in_parts).  Batch
    sizes that do not divide evenly (389 -> 129 + 130 + 130, 257 -> 128 + 129) must give, tile for tile, what calls
    small enough to stay one part give."""
    import torch
    from classpose_b200 import synth
    from classpose_b200.engine import get_engine
    eng = get_engine()
    d = synth.make_batch(389, 256, 256, 7, seed=31)
    for B in (389, 257):
        masks, counts, cc, _ = eng.compute_masks_batch(d["dP"][:B], d["cellprob"][:B], d["logits"][:B])
        for lo in range(0, B, 100):
            hi = min(B, lo + 100)
            m1, c1, cc1, _ = eng.compute_masks_batch(d["dP"][lo:hi], d["cellprob"][lo:hi], d["logits"][lo:hi])
            assert torch.equal(m1, masks[lo:hi]) and torch.equal(c1, counts[lo:hi])
            used = torch.arange(cc1.shape[1], device=cc1.device).view(1, -1) <= c1.view(-1, 1)     # entries 0..count
            assert torch.equal(cc1[used], cc[lo:hi][used])
    assert int(counts.min()) > 0


def test_full_size_batch_properties():
    """BASELINE configs[1] size (1024 conic tiles): size-independent properties of the result.
    labels contiguous 1..counts[b]; every instance within [min_size, 0.4*N]; classes in range; a tile gives the
    bit-identical result alone and inside the batch; the run is deterministic; hole fill is idempotent."""
    import torch
    from classpose_b200 import synth
    from classpose_b200.engine import get_engine
    eng = get_engine()
    B, H, W, C = 1024, 256, 256, 7
    d = synth.make_batch(B, H, W, C, seed=77)
    masks, counts, cc, _ = eng.compute_masks_batch(d["dP"], d["cellprob"], d["logits"])
    masks2, counts2, cc2, _ = eng.compute_masks_batch(d["dP"], d["cellprob"], d["logits"])
    assert torch.equal(masks, masks2) and torch.equal(counts, counts2) and torch.equal(cc, cc2)
    LC = eng.label_capacity(H, W)
    key = (torch.arange(B, device=masks.device).view(B, 1, 1) * LC + masks).reshape(-1).long()
    area = torch.bincount(key, minlength=B * LC).view(B, LC)
    present = area[:, 1:] > 0
    assert torch.equal(present.sum(1).int(), counts)                       # no gaps, no extra labels
    assert torch.equal(masks.view(B, -1).max(1).values.int(), counts)
    lab_area = area[:, 1:][present]
    assert int(lab_area.min()) >= 15 and int(lab_area.max()) <= 0.4 * H * W
    idx = torch.arange(LC, device=masks.device).view(1, LC)
    live = (idx >= 1) & (idx <= counts.view(B, 1))
    assert int(cc[live].min()) >= 0 and int(cc[live].max()) < C
    assert int(counts.sum()) > 80 * B                                      # ~90 planted cells per tile
    for b in (0, 517, 1023):
        m1, c1, k1, _ = eng.compute_masks_batch(d["dP"][b:b + 1], d["cellprob"][b:b + 1], d["logits"][b:b + 1])
        assert torch.equal(m1[0], masks[b]) and int(c1[0]) == int(counts[b])
        assert torch.equal(k1[0, :int(c1[0]) + 1], cc[b, :int(c1[0]) + 1])
    again, cnt_again = eng.fill_holes_and_remove_small_masks(masks[:64], LC, 15)
    assert torch.equal(again, masks[:64]) and torch.equal(cnt_again, counts[:64])
    offs, total = eng.label_offsets(counts, 0)
    assert int(total[0]) == int(counts.sum()) and int(offs[-1]) == int(counts[:-1].sum())


def test_postprocessor_cells_match_reference_loop():
    """Next row N1: the device cell table against the reference PostProcessor loop restated with its own library
    calls (find_objects + cv2.findContours; shapely is absent, its measures are the shoelace formulas)."""
    import cv2
    from scipy.ndimage import find_objects
    from classpose_b200 import postprocess
    from classpose_b200.engine import get_engine
    eng = get_engine()
    tiles = [pc.std_tile(1), pc.adv_tile()]
    dP = np.stack([t["dP"] for t in tiles]); cp = np.stack([t["cellprob"] for t in tiles])
    lg = np.stack([t["logits"] for t in tiles])
    masks, counts, cc, _ = eng.compute_masks_batch(dP, cp, lg)
    table = postprocess.cell_table(masks, counts)
    coords, scale = [(1000, 2000), (1256, 2000)], 1.136
    cells, n_invalid = postprocess.cells_as_reference_dicts(table, cc.cpu().numpy(), coords, scale,
                                                            labels=[f"c{i}" for i in range(7)])
    mh = masks.cpu().numpy()
    for b in range(2):
        ref_cells = []
        for l, slc in enumerate(find_objects(mh[b]), start=1):
            if slc is None:
                continue
            ys, xs = slc
            cell = mh[b][ys, xs] == l
            cs = cv2.findContours(np.uint8(cell), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
            pts = (cs[0][:, 0] + np.array([xs.start, ys.start])) * scale + np.array(coords[b])
            if pts.shape[0] < 4 or not pc.ring_is_simple(cs[0][:, 0]):
                continue
            x, y = pts[:, 0], pts[:, 1]
            xn, yn = np.roll(x, -1), np.roll(y, -1)
            cr = x * yn - xn * y
            ref_cells.append(dict(coords=pts.tolist() + [pts[0].tolist()], area=abs(cr.sum()) / 2,
                                  perimeter=np.hypot(xn - x, yn - y).sum(),
                                  centroid=[((x + xn) * cr).sum() / (3 * cr.sum()), ((y + yn) * cr).sum() / (3 * cr.sum())],
                                  class_int=int(cc[b, l]) - 1))
        assert len(ref_cells) == len(cells[b]) > 0
        for r, c in zip(ref_cells, cells[b]):
            np.testing.assert_allclose(c["coords"], r["coords"], rtol=0, atol=1e-9)
            np.testing.assert_allclose(c["area"], r["area"], rtol=1e-9)
            np.testing.assert_allclose(c["perimeter"], r["perimeter"], rtol=1e-9)
            np.testing.assert_allclose(c["centroid"], np.round(r["centroid"], 2), atol=0.011)
            assert c["class_int"] == r["class_int"]


def test_device_resident_eval_tail_matches_host_path():
    """Next row N3: network sub-tile outputs on the GPU -> blend -> masks -> classes without leaving the device,
    against the reference's host sequence (unaugment, average_tiles, crop, dynamics, class vote) on the oracle."""
    import torch
    from classpose_b200 import core
    # conic (C = 7) and the monusac class count (C = 5) of BASELINE configs[3], each with and without --tta
    for t, augment in ((pc.std_tile(4), False), (pc.std_tile(4), True), (pc.std_tile(23, C=5), True), (pc.std_tile(23, C=5), False)):
        C = t["logits"].shape[0]
        pads, geo = core.tile_layout(256, 256, 256, augment=augment)
        Ly, Lx = geo["Ly"], geo["Lx"]
        full = np.zeros((C + 3, Ly, Lx), np.float32)
        full[:C, pads[0]:pads[0] + 256, pads[2]:pads[2] + 256] = t["logits"]
        full[C:C + 2, pads[0]:pads[0] + 256, pads[2]:pads[2] + 256] = t["dP"]
        full[C + 2, pads[0]:pads[0] + 256, pads[2]:pads[2] + 256] = t["cellprob"]
        # what the network would emit per sub-tile: flipped inputs give flipped outputs with dY / dX sign changes
        tiles = np.zeros((len(geo["y0"]), C + 3, 256, 256), np.float32)
        for j, (y0, x0, f) in enumerate(zip(geo["y0"], geo["x0"], geo["flip"])):
            s = full[:, y0:y0 + 256, x0:x0 + 256].copy()
            if f & 1:
                s = s[:, ::-1]; s[C] *= -1
            if f & 2:
                s = s[:, :, ::-1]; s[C + 1] *= -1
            tiles[j] = s
        # reference host sequence
        ysub = [[a, a + 256] for a in geo["y0"]]; xsub = [[a, a + 256] for a in geo["x0"]]
        yfl = tiles[:, C:].reshape(geo["ny"], geo["nx"], 3, 256, 256).copy()
        ycl = tiles[:, :C].reshape(geo["ny"], geo["nx"], C, 256, 256).copy()
        if augment:
            yfl = otf.unaugment_tiles(yfl); ycl = classpose_ref.unaugment_class_tiles(ycl)
        yf = otf.average_tiles(yfl.reshape(-1, 3, 256, 256), ysub, xsub, Ly, Lx)[:, pads[0]:Ly - pads[1], pads[2]:Lx - pads[3]]
        yc = otf.average_tiles(ycl.reshape(-1, C, 256, 256), ysub, xsub, Ly, Lx)[:, pads[0]:Ly - pads[1], pads[2]:Lx - pads[3]]
        ref = odyn.resize_and_compute_masks(yf[:2], yf[2])
        ref_cm, _ = classpose_ref.compute_class_masks(ref, yc[:, None])
        masks, counts, cc, cm, dP, cellprob = core.eval_tail(torch.from_numpy(tiles[None]).cuda(), C, pads, geo,
                                                              augment=augment, want_class_masks=True)
        # (the float32 error-free blend equals numpy's float64 accumulate except where numpy's double rounding bites:
        #  ~1e-6 of the elements, by one ulp of the accumulator)
        np.testing.assert_allclose(dP[0].cpu().numpy(), yf[:2], rtol=1e-6, atol=1e-6)
        assert np.mean(dP[0].cpu().numpy() == yf[:2]) > 0.9999
        r = metrics.class_agreement(ref, ref_cm, masks[0].cpu().numpy(), cm[0].cpu().numpy().astype(np.int64))
        assert r["f1"] >= 0.995 and not r["class_mismatch"] and r["n_pred"] == r["n_true"]


def test_follow_flows_against_torch_cuda_grid_sample():
    """Independent cross-check on the B200: cellpose's steps_interp op sequence run with torch's CUDA grid_sample (an ATen
    kernel this repo does not own) against k_follow_pool.  float32 rounding differs between implementations and is
    amplified by pixels orbiting a sink, so the bar is the one used against torch-CPU: >= 99.9 % identical truncated end
    points; with few steps nothing can amplify and the match must be (almost) total."""
    import torch
    from classpose_b200.engine import get_engine
    eng = get_engine()
    dev = eng.device
    tiles = [pc.std_tile(s) for s in (1, 3, 4)] + [pc.adv_tile()]
    for niter, bar in ((200, 0.999), (3, 0.9999)):
        tot = same = 0
        for t in tiles:
            H, W = t["cellprob"].shape
            fg = t["cellprob"] > 0
            ys, xs = np.nonzero(fg)
            d = (t["dP"] * fg / 5.0).astype(np.float32)
            pt = torch.zeros((1, 1, len(ys), 2), dtype=torch.float32, device=dev)
            im = torch.zeros((1, 2, H, W), dtype=torch.float32, device=dev)
            pt[0, 0, :, 0] = torch.from_numpy(xs).to(dev).float(); pt[0, 0, :, 1] = torch.from_numpy(ys).to(dev).float()
            im[0, 0] = torch.from_numpy(d[1]).to(dev); im[0, 1] = torch.from_numpy(d[0]).to(dev)
            shape = np.array([W, H]).astype("float") - 1
            for k in range(2):
                im[:, k] *= 2.0 / shape[k]
                pt[..., k] /= shape[k]
            pt *= 2; pt -= 1
            for _ in range(niter):
                dPt = torch.nn.functional.grid_sample(im, pt, align_corners=False)
                for k in range(2):
                    pt[..., k] = torch.clamp(pt[..., k] + dPt[:, k], -1.0, 1.0)
            pt += 1; pt *= 0.5
            for k in range(2):
                pt[..., k] *= shape[k]
            ref = pt[0, 0].int().cpu().numpy()                         # (x, y) truncated
            pf, _ = eng.follow_flows(t["dP"][None], t["cellprob"][None], niter, 0.0)
            pf = pf[0].cpu().numpy()
            eq = ((pf[ys, xs] >> 16) == ref[:, 1]) & ((pf[ys, xs] & 0xFFFF) == ref[:, 0])
            tot += len(ys); same += int(eq.sum())
        assert same / tot >= bar, (niter, same, tot)


def test_dedup_feature_list_signature():
    """Next row N2: same call shape as the reference's deduplicate(features, max_dist)."""
    from classpose_b200 import dedup
    from oracle import dedup as odedup
    centers, sizes = pc.overlap_duplicates(3, n_cells=500, extent=1500.0)
    feats = [{"id": i, "properties": {"measurements": [{"name": "area", "value": float(s)},
                                                       {"name": "centroidX", "value": float(c[0])},
                                                       {"name": "centroidY", "value": float(c[1])}]}}
             for i, (c, s) in enumerate(zip(centers, sizes))]
    out = dedup.deduplicate(feats)
    ref = odedup.reference_greedy(centers, sizes, 7.5)
    assert [f["id"] for f in out] == list(np.nonzero(ref)[0])
    assert dedup.deduplicate([]) == []


def test_dense_512_tile_against_oracle():
    """BASELINE configs[4] shape: one 512x512 tile with ~1.8k small nuclei, all stages on, against the oracle."""
    from classpose_b200.engine import get_engine
    from oracle import synth as osynth
    eng = get_engine()
    t = osynth.make_tile(21, H=512, W=512, C=7, n_grid=45, axes=(3.5, 5.0))
    ref = odyn.resize_and_compute_masks(t["dP"], t["cellprob"])
    ref_cm, _ = classpose_ref.compute_class_masks(ref, t["logits"][:, None])
    masks, counts, cc, cm = eng.compute_masks_batch(t["dP"][None], t["cellprob"][None], t["logits"][None],
                                                    want_class_masks=True)
    r = metrics.class_agreement(ref, ref_cm, masks[0].cpu().numpy(), cm[0].cpu().numpy().astype(np.int64))
    assert r["n_true"] > 1500
    assert r["f1"] >= 0.995 and not r["class_mismatch"], {k: v for k, v in r.items() if k not in ("pairs",)}
    assert abs(r["n_pred"] - r["n_true"]) <= max(1, 0.001 * r["n_true"])


def test_large_and_rectangular_tiles_run():
    """Shapes beyond the benchmark: 1024x768 (large labels take the block-level diffusion / hole-fill paths)."""
    import torch
    from classpose_b200 import synth
    from classpose_b200.engine import get_engine
    eng = get_engine()
    d = synth.make_batch(2, 768, 1024, 5, n_grid=8, axes=(20.0, 40.0), seed=9)       # nuclei up to 80 px across
    masks, counts, cc, _ = eng.compute_masks_batch(d["dP"], d["cellprob"], d["logits"])
    torch.cuda.synchronize()
    planted = torch.stack([(d["labels"][b].unique() > 0).sum() for b in range(2)]).cpu()
    assert (counts.cpu() - planted).abs().max() <= 2
    ref = odyn.resize_and_compute_masks(d["dP"][0].cpu().numpy(), d["cellprob"][0].cpu().numpy())
    assert metrics.match_instances(ref, masks[0].cpu().numpy())["f1"] >= 0.99


def test_prepare_tiles_host_mirror():
    """Next row N4 through the python mirror: uint8 RGB tile -> network input tiles, against the numpy sequence."""
    from classpose_b200 import core
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, size=(2, 225, 225, 3)).astype(np.uint8)
    for augment in (False, True):
        tiles, pads, geo = core.prepare_tiles(img, 256, augment=augment)
        for b in range(2):
            ref, ysub, xsub, rpads = otf.prepare_tiles(img[b].astype(np.float32), 256, augment=augment)
            assert tuple(rpads) == tuple(pads)
            np.testing.assert_array_equal(tiles[b].cpu().numpy(), ref)


def test_batch_slicing_beyond_the_pixel_index_limit():
    """Engine slices batches that exceed one device call's 31-bit pixel index (forced here with a small limit)."""
    import torch
    from classpose_b200.engine import get_engine
    eng = get_engine()
    tiles = [pc.std_tile(s) for s in (1, 3, 4, 6)]
    dP = np.stack([t["dP"] for t in tiles]); cp = np.stack([t["cellprob"] for t in tiles]); lg = np.stack([t["logits"] for t in tiles])
    a = eng.compute_masks_batch(dP, cp, lg, want_class_masks=True)
    b = eng.compute_masks_batch(dP, cp, lg, want_class_masks=True, max_pixels_per_call=256 * 256 + 5)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
