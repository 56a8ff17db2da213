"""Self-consistency of the (unpinned) Cellpose restatement: it must recover planted cells,
and its building blocks must agree with independent formulations."""
import numpy as np
import torch

from oracle import dynamics, metrics, synth, transforms, utils


def test_recovers_planted_cells():
    tile = synth.make_tile(3)
    m = dynamics.resize_and_compute_masks(tile["dP"], tile["cellprob"])
    r = metrics.match_instances(tile["labels"], m)
    assert r["f1"] >= 0.99
    assert m.dtype == np.uint16


def test_no_foreground_returns_zeros():
    dP = np.zeros((2, 32, 32), np.float32)
    m = dynamics.resize_and_compute_masks(dP, -np.ones((32, 32), np.float32))
    assert m.shape == (32, 32) and not m.any()


def test_euler_step_closed_form_matches_grid_sample():
    """Pixel-space form of one step: sx = x*W/(W-1) - 0.5, zero padding."""
    rng = np.random.default_rng(0)
    H, W = 37, 53
    dP = rng.normal(size=(2, H, W)).astype(np.float32)
    ys, xs = np.nonzero(np.ones((H, W), bool))
    p = dynamics.steps_interp(dP, (ys, xs), 1).numpy()

    def bil(f, sy, sx):
        y0, x0 = np.floor(sy).astype(int), np.floor(sx).astype(int)
        wy, wx = sy - y0, sx - x0
        out = np.zeros_like(sy)
        for dy, dx, w in ((0, 0, (1 - wy) * (1 - wx)), (0, 1, (1 - wy) * wx),
                          (1, 0, wy * (1 - wx)), (1, 1, wy * wx)):
            yy, xx = y0 + dy, x0 + dx
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            out[ok] += f[yy[ok], xx[ok]] * w[ok]
        return out

    sy = ys * H / (H - 1) - 0.5
    sx = xs * W / (W - 1) - 0.5
    ey = np.clip(ys + bil(dP[0].astype(np.float64), sy, sx), 0, H - 1)
    ex = np.clip(xs + bil(dP[1].astype(np.float64), sy, sx), 0, W - 1)
    assert np.abs(p[0] - ey).max() < 1e-4 and np.abs(p[1] - ex).max() < 1e-4


def test_renumber_first_appearance():
    a = np.array([[0, 7, 7], [3, 0, 9], [9, 3, 1]], np.uint16)
    np.testing.assert_array_equal(utils.renumber(a), [[0, 1, 1], [2, 0, 3], [3, 2, 4]])


def test_fill_holes_overwrites_enclosed_label_and_size_quirk():
    lab = np.zeros((24, 24), np.int32)
    yy, xx = np.mgrid[0:24, 0:24]
    ring = ((yy - 12) ** 2 + (xx - 12) ** 2 <= 100) & ((yy - 12) ** 2 + (xx - 12) ** 2 > 36)
    lab[ring] = 1
    lab[(yy - 12) ** 2 + (xx - 12) ** 2 <= 9] = 2
    out = utils.fill_holes_and_remove_small_masks(lab, min_size=15)
    assert set(np.unique(out)) == {0, 1}
    assert out[12, 12] == 1 and (out > 0).sum() == ((yy - 12) ** 2 + (xx - 12) ** 2 <= 100).sum()


def test_average_tiles_partition_of_unity():
    rng = np.random.default_rng(1)
    img = rng.normal(size=(3, 272, 272)).astype(np.float32)
    IMG, ysub, xsub, Ly, Lx = transforms.make_tiles(img, bsize=256, augment=False, tile_overlap=0.1)
    assert IMG.shape[:2] == (2, 2)
    y = IMG.reshape(4, 3, 256, 256)
    out = transforms.average_tiles(y, ysub, xsub, Ly, Lx)
    np.testing.assert_allclose(out, img, rtol=0, atol=2e-6)
    # augmented tiles come back after un-augmenting (flows change sign on flips, so feed |.|-free data)
    IMG, ysub, xsub, Ly, Lx = transforms.make_tiles(img, bsize=256, augment=True)
    assert IMG.shape[:2] == (3, 3)
    y = IMG.copy()
    y[:, :, 0] *= 1  # undo sign handling by re-applying it to the data the net would see
    un = transforms.unaugment_tiles(y.copy())
    # channel 2 (cellprob) is never negated
    out = transforms.average_tiles(un.reshape(9, 3, 256, 256), ysub, xsub, Ly, Lx)
    np.testing.assert_allclose(out[2], img[2], rtol=0, atol=2e-6)


def test_pad_geometry_of_a_wsi_tile():
    assert transforms.get_pad_yx(256, 256, min_size=(256, 256)) == (8, 8, 8, 8)
