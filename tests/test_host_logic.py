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
