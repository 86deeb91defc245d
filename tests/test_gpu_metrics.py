"""Per-pose Procrustes metrics on the GPU (gem_pose_align_errors) against the values the reference
printed for the golden end-to-end run and against the host numpy implementation.  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_metrics_match_reference(golden_dir):
    from globalegomocap_b200.metrics import calculate_errors
    g = np.load(os.path.join(golden_dir, "main_mi3.npz"))
    res = calculate_errors(g["final_estimated_seq"], g["final_estimated_seq"], g["final_optimized_seq"], g["final_gt_seq"],
                           on_device=True)
    for k in ("aligned_original_mpjpe", "aligned_optimized_mpjpe", "bone_length_aligned_original_mpjpe",
              "bone_length_aligned_optimized_mpjpe", "joints_error", "original_global_mpjpe",
              "optimized_aligned_global_mpjpe"):
        np.testing.assert_allclose(res[k], g["err__" + k], rtol=1e-9, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("n", [1, 7, 1000])
def test_device_alignment_matches_host(n):
    from globalegomocap_b200 import metrics
    rng = np.random.default_rng(n)
    gt = rng.uniform(-0.8, 0.8, (n, 15, 3)) + np.array([0.0, 0.0, 1.0])
    est = gt + rng.normal(0, 0.05, gt.shape)
    est[0] = gt[0] @ np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]]) * 1.7 + 0.3        # an exact similarity: error 0
    if n > 2:
        est[1] = gt[1] * np.array([1.0, 1.0, -1.0])                                    # a reflection (det fix branch)
        est[2, :, 2] = 0.25                                                             # a planar pose (rank-2 covariance)
    for resize in (False, True):
        a_h, g_h = metrics._align_per_pose_host(est, gt, resize)
        a_d, g_d = metrics._align_per_pose_device(est, gt, resize)
        assert np.abs(g_d - g_h).max() < 1e-12, ("gt", resize, float(np.abs(g_d - g_h).max()))
        worst = np.abs(a_d - a_h).reshape(n, -1).max(axis=1)
        assert worst.max() < 1e-9, ("aligned", resize, int(worst.argmax()), float(worst.max()))
    a_d, _ = metrics._align_per_pose_device(est, gt, False)
    assert np.abs(a_d[0] - gt[0]).max() < 1e-12, float(np.abs(a_d[0] - gt[0]).max())
