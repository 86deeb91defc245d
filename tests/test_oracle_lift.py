"""The oracle's initial 3-D lift (oracle/lift_np.py) against vectors recorded from the unmodified reference
(tests/golden/make_golden_lift.py -> lift.npz): SURVEY.md §8f N3."""
import os

import numpy as np
import pytest

from oracle import lift_np


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "lift.npz"))


def test_max_preds_match_reference(gold):
    preds, maxvals, idx = lift_np.max_preds(gold["heat"])
    assert preds.dtype == np.float32
    assert np.array_equal(preds, gold["preds"])                 # integer pixel coordinates: exact
    assert np.array_equal(maxvals, gold["maxvals"])
    # edge cases built into frame 1: first row-major maximum wins ties, borders, non-positive maps -> (0, 0)
    assert tuple(preds[1, 0]) == (7 * 16 + 128, 5 * 16)
    assert tuple(preds[1, 1]) == (128.0, 0.0) and tuple(preds[1, 2]) == (63 * 16 + 128, 63 * 16)
    assert tuple(preds[1, 6]) == (63 * 16 + 128, 0.0)
    for k in (3, 4, 5):
        assert tuple(preds[1, k]) == (0.0, 0.0) and maxvals[1, k] == 0.0


def test_lift_matches_reference(gold):
    pts, _, _, _ = lift_np.lift(gold["heat"], gold["depth"], gold["center"], gold["poly_c2w"])
    assert np.abs(pts - gold["points"]).max() < 1e-13
    # every lifted joint sits at its depth from the camera
    assert np.abs(np.linalg.norm(pts, axis=-1) - gold["depth"]).max() < 1e-12
    res, _, _, _ = lift_np.lift(gold["heat"], gold["depth"], gold["center"], gold["poly_c2w"], gold["bone_length"])
    assert np.abs(res - gold["resized"]).max() < 1e-12
