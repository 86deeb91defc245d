"""Initial 3-D lift kernel (gem_lift_skeleton) through the C ABI against vectors recorded from the
unmodified reference and against the oracle.  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "lift.npz"))


def test_lift_matches_reference_vectors(gold):
    from globalegomocap_b200.lift import lift_skeletons
    pts, preds, maxvals = lift_skeletons(gold["heat"], gold["depth"], gold["poly_c2w"], gold["center"][0], gold["center"][1],
                                         return_preds=True)
    assert np.array_equal(preds, gold["preds"])              # pixel coordinates: exact (ties, borders, masked maps)
    assert np.array_equal(maxvals, gold["maxvals"])
    assert np.abs(pts - gold["points"]).max() < 1e-13         # float64 back-projection
    res = lift_skeletons(gold["heat"], gold["depth"], gold["poly_c2w"], gold["center"][0], gold["center"][1],
                         bone_length=gold["bone_length"])
    assert np.abs(res - gold["resized"]).max() < 1e-12


@pytest.mark.parametrize("n", [1, 2, 37, 300])
def test_lift_matches_oracle_on_random_maps(gold, n):
    from globalegomocap_b200.lift import lift_skeletons
    from oracle import lift_np
    rng = np.random.default_rng(100 + n)
    heat = rng.standard_normal((n, 64, 64, 15)).astype(np.float32)
    heat[:, ::7, ::5, :] = np.round(heat[:, ::7, ::5, :] * 2) / 2          # plenty of exact ties
    if n > 1:
        heat[1, :, :, 4] = -1.0                                            # a masked map
    depth = rng.uniform(0.2, 3.0, (n, 15))
    ref_pts, ref_preds, ref_max, _ = lift_np.lift(heat, depth, gold["center"], gold["poly_c2w"])
    pts, preds, maxvals = lift_skeletons(heat, depth, gold["poly_c2w"], gold["center"][0], gold["center"][1],
                                         return_preds=True)
    assert np.array_equal(preds, ref_preds) and np.array_equal(maxvals, ref_max)
    assert np.abs(pts - ref_pts).max() < 1e-13


def test_lift_rejects_unsupported_shapes(gold):
    from globalegomocap_b200._lib import GemError
    from globalegomocap_b200.lift import lift_skeletons
    with pytest.raises(GemError):
        lift_skeletons(np.zeros((1, 10, 10, 15), np.float32), np.ones((1, 15)), gold["poly_c2w"], 1.0, 1.0)
