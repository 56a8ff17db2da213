"""Mutation tests: every constant of the Cellpose op sequence that SURVEY.md marks [RECALLED] (seed threshold 10, growth
threshold 2, five growth steps, window padding rpad, flow scale 1/5, no log before the diffusion gradient, the 0.4 /
15-pixel defaults are arguments) is changed in the ORACLE, one at a time, and the parity comparison with the kernels
(compiled unchanged against the CPU grid simulator) must then FAIL.  This proves the parity suite is sensitive to each of
them: the day the real cellpose source contradicts one, changing the oracle constant turns the suite red until the
kernel follows."""
import numpy as np
import pytest

import parity_cases as pc
from backends import SimBackend
from oracle import dynamics


@pytest.fixture(scope="module")
def be():
    return SimBackend()


def _get_masks_inputs():
    """Hand-made end points with bins of exactly 10, 11 and 12 points, windows that overlap, bins of 2 / 3 points around a
    seed (growth threshold) and chains longer than five growth steps."""
    H, W = 64, 80
    ys, xs = np.nonzero(np.ones((H, W), bool))
    pf = np.zeros((2, H * W), np.int32)
    targets = [(8, 8, 10), (8, 30, 11), (8, 50, 12), (30, 20, 40), (30, 60, 25), (50, 40, 30)]
    i = 0
    for y, x, n in targets:
        pf[0, i:i + n] = y; pf[1, i:i + n] = x; i += n
    # halo around (30, 20): rings of 3-point and 2-point bins out to distance 7
    for d in range(1, 8):
        for (yy, xx) in ((30, 20 + d), (30, 20 - d), (30 + d, 20)):
            n = 3 if d % 2 else 2
            pf[0, i:i + n] = yy; pf[1, i:i + n] = xx; i += n
    # a line of 3-point bins 8 long from (50, 40): only five growth steps reach
    for d in range(1, 9):
        pf[0, i:i + 3] = 50; pf[1, i:i + 3] = 40 + d; i += 3
    # a hook inside the 11 x 11 window of (50, 40) whose last bin is six growth steps away along the chain
    for (yy, xx) in ((51, 43), (52, 43), (53, 42), (53, 41)):
        pf[0, i:i + 3] = yy; pf[1, i:i + 3] = xx; i += 3
    # everything else parks on far, sub-threshold bins (one point each)
    rest = np.arange(i, H * W)
    pf[0, rest] = 60 + (rest % 3); pf[1, rest] = (rest * 7) % W
    return pf, (ys, xs), (H, W)


def _kernel_get_masks(be):
    pf, inds, (H, W) = _get_masks_inputs()
    packed = np.full((H, W), -1, np.int32)
    packed[inds] = (pf[0] << 16) | pf[1]
    m, _ = be.get_masks(packed[None], 0.4)
    return m[0], pf, inds, (H, W)


def test_unmutated_oracle_agrees(be):
    m, pf, inds, shape = _kernel_get_masks(be)
    np.testing.assert_array_equal(m, dynamics.get_masks(pf, inds, shape))
    pc.case_follow_flows_few_iters_exact(be)
    pc.case_masks_to_flows_exact(be)


@pytest.mark.parametrize("name,value", [("SEED_MIN", 9), ("SEED_MIN", 11), ("GROW_MIN", 1), ("GROW_MIN", 3),
                                        ("GROW_ITERS", 4), ("GROW_ITERS", 6), ("RPAD", 3)])
def test_get_masks_constants_are_pinned_by_the_suite(be, monkeypatch, name, value):
    m, pf, inds, shape = _kernel_get_masks(be)
    monkeypatch.setattr(dynamics, name, value)
    try:
        ref = dynamics.get_masks(pf, inds, shape, rpad=dynamics.RPAD)
    except Exception:
        return              # the mutated oracle does not even run (rpad smaller than the window): mutation detected
    assert not np.array_equal(m, ref), f"{name}={value} went unnoticed"


def test_flow_scale_is_pinned_by_the_suite(be, monkeypatch):
    monkeypatch.setattr(dynamics, "FLOW_SCALE", 4.0)
    pc._tile_cache.clear()
    try:
        with pytest.raises(AssertionError):
            pc.case_remove_bad_flow_masks_exact(be)        # flow error uses dP / 5
        t = pc.std_tile(1)
        fg = t["cellprob"] > 0
        p = dynamics.follow_flows(t["dP"] * fg / dynamics.FLOW_SCALE, np.nonzero(fg), 3).int().numpy()
        pf, _ = be.follow_flows(pc.f32(t["dP"][None]), pc.f32(t["cellprob"][None]), 3, 0.0)
        ys, xs = np.nonzero(fg)
        eq = ((pf[0][ys, xs] >> 16) == p[0]) & ((pf[0][ys, xs] & 0xFFFF) == p[1])
        assert eq.mean() < 0.99, "follow_flows does not notice the flow scale"
    finally:
        pc._tile_cache.clear()


def test_no_log_before_gradient_is_pinned_by_the_suite(be, monkeypatch):
    monkeypatch.setattr(dynamics, "DIFFUSE_LOG", True)
    with pytest.raises(AssertionError):
        pc.case_masks_to_flows_exact(be)


def test_neighbour_order_and_ninth_are_pinned(be, monkeypatch):
    """masks_to_flows agrees to 1e-12; a different averaging constant (8 instead of 9 neighbours) must not."""
    orig = dynamics.extend_centers

    def mutated(neighbors, centers, isneighbor, shape, n_iter):
        return orig(neighbors[:, :8], centers, isneighbor[:8], shape, n_iter)
    monkeypatch.setattr(dynamics, "extend_centers", mutated)
    with pytest.raises(Exception):
        pc.case_masks_to_flows_exact(be)
