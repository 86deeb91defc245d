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


def test_load_clips_lands_every_clip_in_one_staging_buffer(tmp_path):
    """optimizer.load_clips (reference optimizer.py:315-324): the pickles' per-frame lists are stacked into one buffer
    per key (pinned when there is a GPU), the clips are views of it, extra keys are ignored, dtypes are the
    reference's.  Without CUDA the same code runs on pageable memory."""
    import pickle

    from globalegomocap_b200 import optimizer as gem
    from globalegomocap_b200 import synthetic as syn
    clips = [syn.make_clip(n, seed=40 + i) for i, n in enumerate((18, 26))]
    dirs = []
    for i, c in enumerate(clips):
        d = tmp_path / f"data_start_{i}_end_{i + 1}"
        syn.write_clip_pickle(c, str(d))
        with open(d / "test_data.pkl", "rb") as f:          # an extra key like the data-prep script writes
            raw = pickle.load(f)
        raw["estimated_global_skeleton"] = raw["gt_global_skeleton"]
        with open(d / "test_data.pkl", "wb") as f:
            pickle.dump(raw, f)
        dirs.append(str(d))
    assert gem.load_clips(dirs).planar == 2                  # default: tiled (64 x 64 maps tile)
    planar = gem.load_clips(dirs, planar=True)               # heat maps land as [frames, J, H, W]
    assert planar.planar == 1 and tuple(planar.heat_all.shape) == (18 + 26, 15, 64, 64)
    for c, got in zip(clips, planar):
        assert np.array_equal(got["heatmap_list"].numpy(), c["heatmap_list"].transpose(0, 3, 1, 2))
        assert np.array_equal(got["gt_global_skeleton"].numpy(), c["gt_global_skeleton"])
    tiled = gem.load_clips(dirs, planar="tiled")             # [frames, J, H/4, W/8, 4, 8]: 4 x 8 tiles of every map
    assert tiled.planar == 2 and tuple(tiled.heat_all.shape) == (18 + 26, 15, 16, 8, 4, 8)
    from globalegomocap_b200.engine import heat_dims, untile_heat
    assert heat_dims(tiled.heat_all.shape, 2) == (64, 64, 15)
    for c, got in zip(clips, tiled):
        assert np.array_equal(untile_heat(got["heatmap_list"]).numpy(), c["heatmap_list"].transpose(0, 3, 1, 2))
        y, x = 37, 29                                        # a texel's address inside its map
        flat = got["heatmap_list"].reshape(len(c["heatmap_list"]), 15, -1).numpy()
        assert np.array_equal(flat[:, :, ((y >> 2) * 8 + (x >> 3)) * 32 + (y & 3) * 8 + (x & 7)], c["heatmap_list"][:, y, x, :])
    cs = gem.load_clips(dirs, planar=False)
    assert isinstance(cs, gem.ClipSet) and len(cs) == 2 and not cs.planar
    assert tuple(cs.heat_all.shape) == (18 + 26, 64, 64, 15) and cs.heat_all.dtype == torch.float32
    off = 0
    for c, got in zip(clips, cs):
        assert set(got) == set(gem.CLIP_KEYS)
        for k in gem.CLIP_KEYS:
            assert np.array_equal(got[k].numpy(), c[k]), k
        n = len(c["heatmap_list"])
        assert got["heatmap_list"].data_ptr() == cs.heat_all[off:off + n].data_ptr()      # a view, not a copy
        assert got["estimated_local_skeleton"].dtype == torch.float64
        off += n
    one = gem.load_clip(dirs[1])
    assert np.array_equal(one["camera_pose_list"].numpy(), clips[1]["camera_pose_list"])
    # the maps are stacked piece by piece on a few host threads: ragged pieces give the same buffer
    chunk = gem._STACK_CHUNK
    try:
        gem._STACK_CHUNK = 5
        for layout in ("tiled", True):
            again = gem.load_clips(dirs, planar=layout)      # (the staging pool is reused: compare with the source)
            for c, got in zip(clips, again):
                maps = untile_heat(got["heatmap_list"]) if layout == "tiled" else got["heatmap_list"]
                assert np.array_equal(maps.numpy(), c["heatmap_list"].transpose(0, 3, 1, 2))
    finally:
        gem._STACK_CHUNK = chunk


def test_bench_workload_resolution_and_strong_shards():
    """bench.py --windows tiles BASELINE configs[4]'s sweep points; --scaling strong cuts every sequence's windows
    into contiguous per-rank blocks whose frames cover each window once."""
    import argparse

    import bench
    from globalegomocap_b200 import distributed as gd
    ns = argparse.Namespace(windows=0, sequences=5, frames=3000)
    assert bench.resolve_workload(ns) == (5, 3000, 1)
    for want, got in ((1000, (3, 3000, 1)), (10000, (27, 3000, 1)), (100000, (27, 3000, 10))):
        ns.windows = want
        assert bench.resolve_workload(ns) == got
        n_seq, frames, rep = got
        assert abs(n_seq * 374 * rep - want) / want < 0.15
    world, frames = 4, 250
    n_win = len(range(0, frames - 10 + 1, 8))
    seen = []
    for r in range(world):
        part = bench.make_strong_clips(1, frames, r, world)[0]
        lo, hi = gd.shard_range(n_win, r, world)
        assert len(part["estimated_local_skeleton"]) == 8 * (hi - lo - 1) + 10
        assert part["mean_bone_length"].shape == (15,)
        seen += list(range(lo, hi))
    assert seen == list(range(n_win))
