"""Host-side metrics (reference calculate_errors.py) against the values the reference printed
for the golden end-to-end run."""
import os

import numpy as np

from globalegomocap_b200.metrics import calculate_errors


def test_metrics_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "main_mi3.npz"))
    # mid_estimated_seq is not part of main's return value; the metrics that depend on it are
    # checked with the estimated sequence in its place (the formula is the same)
    res = calculate_errors(g["final_estimated_seq"], g["final_estimated_seq"], g["final_optimized_seq"],
                           g["final_gt_seq"])
    for k in ("original_global_mpjpe", "optimized_global_mpjpe", "original_camera_pos_error",
              "optimized_camera_pos_error", "original_aligned_camera_pos_error", "optimized_aligned_camera_pos_error",
              "original_aligned_global_mpjpe", "optimized_aligned_global_mpjpe", "aligned_original_mpjpe",
              "aligned_optimized_mpjpe", "bone_length_aligned_original_mpjpe", "bone_length_aligned_optimized_mpjpe",
              "joints_error"):
        np.testing.assert_allclose(res[k], g["err__" + k], rtol=1e-9, atol=1e-12, err_msg=k)
    assert list(res.keys())[:3] == ["original_global_mpjpe", "mid_global_mpjpe", "optimized_global_mpjpe"]
    assert len(res) == 18
