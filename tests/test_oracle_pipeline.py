"""oracle/pipeline_np.py pinned against scipy and the reference's optimizer.main
(tests/golden/main_mi3.npz)."""
import os

import numpy as np
from scipy.ndimage import gaussian_filter1d

from oracle import pipeline_np as pl


def test_window_partition_bit_exact():
    assert pl.window_starts(58) == [0, 8, 16, 24, 32, 40, 48]
    assert pl.window_starts(1000)[-1] == 984 and len(pl.window_starts(1000)) == 124
    assert len(pl.window_starts(3000)) == 374 and len(pl.window_starts(100)) == 12
    assert pl.window_starts(9) == [] and pl.window_starts(10) == [0] and pl.window_starts(17) == [0]


def test_gaussian_filter_matches_scipy():
    rng = np.random.default_rng(0)
    for n in (3, 5, 9, 58, 994):
        a = rng.standard_normal((n, 15, 3))
        np.testing.assert_allclose(pl.gaussian_filter1d_reflect(a, 1.0, 0), gaussian_filter1d(a, sigma=1, axis=0),
                                   rtol=0, atol=1e-14)
    k = pl.gaussian_kernel1d(1.0)
    assert len(k) == 9 and abs(k[4] - 0.39894) < 2e-4


def test_merge_lengths_and_overlap_average():
    w = np.arange(3 * 10 * 15 * 3, dtype=np.float64).reshape(3, 10, 15, 3)
    m = pl.merge_batches(w)
    assert m.shape == (8 * 3 + 2, 15, 3)
    np.testing.assert_array_equal(m[:8], w[0, :8])
    np.testing.assert_array_equal(m[8:10], (w[0, 8:] + w[1, :2]) / 2)
    np.testing.assert_array_equal(m[10:16], w[1, 2:8])
    np.testing.assert_array_equal(m[-2:], w[2, 8:])
    assert pl.merge_batches(w[:1]).shape == (10, 15, 3)


def test_main_end_to_end_matches_reference(golden_dir, clip58, vae_weights, camera):
    g = np.load(os.path.join(golden_dir, "main_mi3.npz"))
    out = pl.run_sequence(clip58, vae_weights[0], vae_weights[1], camera, g["eps"], max_iter=int(g["max_iter"]))
    np.testing.assert_allclose(out["est_global"], g["final_estimated_seq"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["gt"], g["final_gt_seq"], rtol=0, atol=1e-12)
    assert out["opt_global"].shape == g["final_optimized_seq"].shape == (58, 15, 3)
    # windows 2 and 3 (frames 16..33) hit a degenerate cubic interpolation in the local stage
    # (d1^2 - g1*g2 changes sign with a 1-ulp change of the loss, see test_oracle_lbfgs.LOOSE)
    ok = np.r_[0:16, 34:58]
    mid_mm = np.abs(out["mid_local"] - g["mid_local_pose_seq"]).max(axis=(1, 2)) * 1000
    assert mid_mm[ok].max() < 0.5, mid_mm
    assert mid_mm.max() < 60
    opt_mm = np.abs(out["opt_global"] - g["final_optimized_seq"]).max(axis=(1, 2)) * 1000
    assert np.median(opt_mm) < 0.5 and opt_mm.max() < 60, opt_mm


def test_stitching_and_slam_transforms_match_the_reference_at_1000_frames(golden_dir):
    """BASELINE configs[0] (tests/golden/main_1k_mi25.npz, the unmodified reference's `optimizer.main` on a 1000-frame
    sequence): the un-optimised estimate goes local -> first camera of its window -> global and through the overlap
    merge (optimizer.py:394-402, 425-437) — the oracle's restatement of that path reproduces the reference's 994
    frames to round-off, and the ground truth passes through the merge unchanged."""
    from globalegomocap_b200 import synthetic as syn
    g = np.load(os.path.join(golden_dir, "main_1k_mi25.npz"))
    clip = syn.make_clip(int(g["n_frames"]), seed=int(g["seed"]))
    starts = pl.window_starts(len(clip["estimated_local_skeleton"]))
    assert len(starts) == int(g["windows"]) == 124
    est_w, gt_w = [], []
    for s in starts:
        cams = clip["camera_pose_list"][s:s + 10]
        rel = pl.relative_global_pose(clip["estimated_local_skeleton"][s:s + 10], cams)
        est_w.append(pl.to_global_pose(rel, cams))
        gt_w.append(clip["gt_global_skeleton"][s:s + 10])
    est = pl.merge_batches(np.stack(est_w))
    assert est.shape == g["final_estimated_seq"].shape == (994, 15, 3)
    np.testing.assert_allclose(est, g["final_estimated_seq"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(pl.merge_batches(np.stack(gt_w)), g["final_gt_seq"])
    # the reference's two runs (eight threads, one thread) of the 25-iteration solve end tens of millimetres apart
    # per frame while their printed metrics agree to 0.13 %: the spread the GPU test measures the CUDA path against
    opt_mm = np.abs(g["final_optimized_seq_one_thread"] - g["final_optimized_seq"]).max(axis=(1, 2)) * 1000
    assert 5.0 < np.median(opt_mm) < 100.0
    worst = max(abs(float(np.mean(g["err1__" + k[5:]])) / float(np.mean(g[k])) - 1.0)
                for k in g.files if k.startswith("err__") and "original" not in k)
    assert worst < 0.005
