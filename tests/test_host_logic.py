"""Host-side bookkeeping that needs no GPU: window partition, upload pieces, bench accounting."""
import numpy as np
import pytest
import torch

from globalegomocap_b200.pipeline import upload_pieces, window_starts


@pytest.mark.parametrize("n_frames", [9, 10, 17, 18, 100, 800, 1000, 1600, 3000, 3001, 3007])
def test_window_starts_match_reference_range(n_frames):
    """range(0, N - 10 + 1, 8) of optimizer.py:370; the trailing (N - 10) mod 8 frames are dropped."""
    assert window_starts(n_frames) == list(range(0, n_frames - 10 + 1, 8))


@pytest.mark.parametrize("n_frames", [10, 100, 800, 1600, 3000, 3005, 12000])
@pytest.mark.parametrize("min_piece", [1, 96, 200])
def test_upload_pieces_cover_the_clip_once_and_hold_their_windows(n_frames, min_piece):
    starts = window_starts(n_frames)
    pieces = upload_pieces(len(starts), n_frames, 10, 2, min_piece)
    assert pieces[0][0] == 0 and pieces[-1][1] == len(starts)
    assert pieces[0][2] == 0 and pieces[-1][3] == n_frames
    for (a0, b0, f0, f1), (a1, b1, g0, g1) in zip(pieces[:-1], pieces[1:]):
        assert b0 == a1 and f1 == g0                    # consecutive, disjoint
        assert a1 % 2 == 0                              # even slice starts
    for a, b, f0, f1 in pieces:
        assert b > a
        assert b - a >= min(min_piece, len(starts)) or len(pieces) == 1
        # every frame a window of this piece reads has landed once this piece (and the earlier ones) has
        assert starts[b - 1] + 10 <= f1


def test_bench_lbfgs_row_accounting():
    import bench
    sol = {s: {"n_iter": torch.tensor([1, 3, 25]), "func_evals": torch.tensor([2, 4, 31])} for s in ("local", "glob")}
    # window 0: 1 iteration (10 rows) + 1 extra evaluation (5); window 1: 3 iterations (30 + 4*1) + 1 evaluation;
    # window 2: 25 iterations (250 + 4*(24*23/2)) + 6 evaluations
    per_stage = (10 + 5) + (30 + 4 + 5) + (250 + 4 * 276 + 30)
    assert bench.lbfgs_rows(sol) == 2 * per_stage


def test_lift_host_pieces_match_reference_vectors(golden_dir):
    """The host-side parts of the initial 3-D lift (calibration loader, bone-length normalisation) against the
    vectors recorded from the unmodified reference; the kernel itself is checked in tests/test_gpu_lift.py."""
    import os

    import numpy as np

    from globalegomocap_b200 import synthetic as syn
    from globalegomocap_b200.lift import load_camera_c2w, skeleton_resize
    g = np.load(os.path.join(golden_dir, "lift.npz"))
    poly, cx, cy = load_camera_c2w(syn.DEFAULT_CAMERA_JSON)
    assert np.array_equal(poly, g["poly_c2w"]) and (cx, cy) == tuple(g["center"])
    for f in range(g["points"].shape[0]):
        out = skeleton_resize(g["points"][f], g["bone_length"])
        assert np.abs(out - g["resized"][f]).max() < 1e-12
    # the input is not modified (the reference's version works in place on its own copy)
    p0 = g["points"][0].copy()
    skeleton_resize(p0, g["bone_length"])
    assert np.array_equal(p0, g["points"][0])
