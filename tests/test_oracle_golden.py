"""Pin the oracle: reference-owned functions against vectors produced by the reference's
own source (tests/golden/make_golden.py), plus the known-answer cases of the reference's
tests/test_remove_border_instances.py:30-117 restated on the same layouts."""
import os

import numpy as np
import pytest

from oracle import classpose_ref as ref


def test_class_vote_matches_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_class_vote.npz"))
    for k in range(int(g["ncases"])):
        cm, uniq = ref.compute_class_masks(g[f"masks{k}"].copy(), g[f"logits{k}"].copy())
        assert cm.dtype == g[f"class_masks{k}"].dtype
        np.testing.assert_array_equal(cm, g[f"class_masks{k}"])
        np.testing.assert_array_equal(uniq, g[f"unique{k}"])


def test_border_removal_matches_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_border.npz"))
    for k in range(int(g["ncases"])):
        for tag in ("2d", "3d"):
            out = ref.remove_border_instances(g[f"in{tag}_{k}"].copy())
            np.testing.assert_array_equal(out, g[f"out{tag}_{k}"])


def test_unaugment_matches_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_unaugment.npz"))
    for k in range(int(g["ncases"])):
        out = ref.unaugment_class_tiles(g[f"in{k}"].copy())
        np.testing.assert_array_equal(out, g[f"out{k}"])


# ---- known-answer cases (layouts as in the reference's test module) --------------------
def six_by_six():
    m = np.zeros((6, 6), np.int64)
    m[0:3, 0:3] = 1
    m[0:3, 3:6] = 2
    m[2:4, 2:4] = 3
    m[3:6, 3:6] = 4
    return m


def test_border_known_answers_2d():
    out = ref.remove_border_instances(six_by_six())
    assert set(np.unique(out)) == {0, 3}
    assert out[2, 2] == 3 and out[2, 3] == 3 and out[3, 2] == 3 and out[3, 3] == 0

    m = np.zeros((4, 4), np.int64); m[0:2] = 1; m[2:4] = 2
    assert not ref.remove_border_instances(m).any()

    assert not ref.remove_border_instances(np.zeros((5, 5), np.int64)).any()

    m = np.zeros((5, 5), np.int64); m[1:4, 1:4] = 7
    out = ref.remove_border_instances(m)
    assert (out[1:4, 1:4] == 7).all() and out[0].sum() == 0 and out[:, 0].sum() == 0


def test_border_known_answers_with_class_channel():
    inst = six_by_six()
    cls = np.zeros_like(inst)
    for i, c in ((1, 1), (2, 2), (3, 3), (4, 1)):
        cls[inst == i] = c
    out = ref.remove_border_instances(np.stack([inst, cls], -1))
    assert set(np.unique(out[..., 0])) == {0, 3} and set(np.unique(out[..., 1])) == {0, 3}
    for (y, x, c), r in (((2, 2, 0), 3), ((2, 3, 0), 3), ((3, 2, 0), 3), ((3, 3, 0), 0),
                         ((2, 2, 1), 3), ((2, 3, 1), 3), ((3, 2, 1), 3), ((3, 3, 1), 0)):
        assert out[y, x, c] == r

    inst = np.zeros((4, 4), np.int64); inst[0:2] = 1; inst[2:4] = 2
    assert not ref.remove_border_instances(np.stack([inst, np.ones_like(inst)], -1)).any()
