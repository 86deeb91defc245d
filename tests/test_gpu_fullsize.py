"""BASELINE configs[2] at its FULL size — 5 synthetic sequences x 3000 frames = 1870 windows, both stages, max_iter 25,
the batch `bench.py` times — checked through properties that do not need the (CPU-hours) reference answer:

  * every window finishes cleanly (status 0, evaluation and iteration counts inside torch's limits
    `optimizer.py:261-262`: max_iter 25, max_eval 31);
  * run-to-run determinism of the whole batch, and independence from how the library slices it (one slice, the default
    four, eight): bit-identical poses and counters;
  * batch independence at scale: windows picked across clips and slice boundaries, solved ALONE from the same inputs
    and noise, reproduce their rows of the 1870-window solve bit for bit (a window is the reference's unit of work,
    `optimizer.py:370-423`: nothing may leak between windows);
  * the stitched sequences have the reference's length 8 W + 2 per clip (`optimizer.py:425-437`), and the SLAM round
    trip local -> first camera of the window -> global (`utils/utils.py:99-112`, `optimizer.py:302-308`), stitched over
    the overlaps, returns camera_pose[f] . x_local[f] for every frame;
  * the 1e4-window sweep point of BASELINE configs[4] (11 220 windows: six replicas of every window in one batch)
    reproduces the 1870-window batch replica by replica, bit for bit.

Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest
import torch

from globalegomocap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

N_SEQ, N_FRAMES = 5, 3000
W_TOTAL = N_SEQ * len(range(0, N_FRAMES - 10 + 1, 8))          # 1870


@pytest.fixture(scope="module")
def full(vae_weights, camera):
    from globalegomocap_b200 import optimizer as gem
    from globalegomocap_b200.engine import Engine
    from globalegomocap_b200.pipeline import WindowBatch
    from globalegomocap_b200.vae_prep import PreparedVae
    clips = [syn.make_clip(N_FRAMES, seed=17 + s) for s in range(N_SEQ)]      # bench.py's rank-0 workload
    eng = Engine(max_windows=W_TOTAL, max_history=24)
    prep = (PreparedVae(vae_weights[0], eng.device), PreparedVae(vae_weights[1], eng.device))
    kw = dict(camera_model_path=camera, final_smooth=False, max_iter=25, local_vae_path=prep[0], global_vae_path=prep[1],
              engine=eng, outputs="all")
    batch = WindowBatch(eng, clips)
    assert batch.W == W_TOTAL == 1870
    eps = torch.randn(W_TOTAL, 2, 2048, generator=torch.Generator().manual_seed(31))
    out = gem.solve_clips(batch, eps=eps, **kw)
    torch.cuda.synchronize()
    yield dict(gem=gem, eng=eng, kw=kw, clips=clips, batch=batch, eps=eps, out=out)
    eng.close()


def _rows(sol):
    return (sol["local"]["pose"], sol["glob"]["pose"], sol["local"]["n_iter"], sol["glob"]["n_iter"],
            sol["local"]["func_evals"], sol["glob"]["func_evals"], sol["local"]["status"])


def test_every_window_finishes_inside_torchs_limits(full):
    loc, glo, it_l, it_g, ev_l, ev_g, status = (t.cpu() for t in _rows(full["out"]["sol"]))
    assert int(status.sum()) == 0
    assert torch.isfinite(loc).all() and torch.isfinite(glo).all()
    for it, ev in ((it_l, ev_l), (it_g, ev_g)):
        assert int(it.min()) >= 1 and int(it.max()) <= 25
        assert int(ev.min()) >= 1 and int(ev.max()) <= 32          # max_eval = 31 line-search evaluations + the first
        assert (ev >= it).all()
    # the optimiser did move every window, and not far: the anchor term ties the pose to the estimate
    x0 = full["batch"].gather(full["batch"].est).to(torch.float32).cpu()
    move = (loc - x0).abs().amax(dim=(1, 2, 3))
    assert float(move.min()) > 0.0 and float(move.max()) < 10.0


def test_full_batch_is_deterministic_and_independent_of_slicing(full):
    gem, eng = full["gem"], full["eng"]
    want = [t.clone() for t in _rows(full["out"]["sol"])]
    try:
        for chunks in (0, 1, 8):                                    # 0 = automatic (the run above)
            eng.set_chunks(chunks)
            again = gem.solve_clips(full["batch"], eps=full["eps"], **full["kw"])
            torch.cuda.synchronize()
            for a, b in zip(_rows(again["sol"]), want):
                assert torch.equal(a, b), chunks
    finally:
        eng.set_chunks(0)


def test_windows_solved_alone_reproduce_their_rows(full):
    from globalegomocap_b200.pipeline import stage_weights
    from globalegomocap_b200.engine import lbfgs_params
    eng, batch, sol = full["eng"], full["batch"], full["out"]["sol"]
    per_clip = W_TOTAL // N_SEQ
    # first / last window of clips, both sides of the automatic slice boundaries, and a random handful
    idx = sorted({0, 1, per_clip - 1, per_clip, 2 * per_clip - 1, 466, 467, 468, 469, 934, 935, 936, 1402, 1403, W_TOTAL - 1}
                 | set(np.random.default_rng(4).integers(0, W_TOTAL, 17).tolist()))
    idx_t = torch.tensor(idx, device=eng.device)
    x_local = batch.gather(batch.est).to(torch.float32)
    cams = batch.gather(batch.cams)
    w_local, w_global = stage_weights(0.0, 0.001, 0.01, 0.01, 0.01)          # solve_clips' defaults (optimize_whole_sequence.py:9-23)
    eng.set_heat_layout(getattr(batch, "planar", False))
    alone = eng.solve_windows(x_local[idx_t], batch.heat, batch.frame_base[idx_t], batch.clip_idx[idx_t], batch.mean_bone,
                              cams[idx_t], full["eps"][idx], w_local, w_global, lbfgs_params(lr=2, max_iter=25))
    torch.cuda.synchronize()
    for a, b in zip(_rows(alone), _rows(sol)):
        assert torch.equal(a, b[idx_t]), [i for i, (p, q) in enumerate(zip(a, b[idx_t])) if not torch.equal(p, q)][:5]


def test_stitched_lengths_and_slam_round_trip(full):
    merged, clips = full["out"]["merged"], full["clips"]
    assert len(merged) == N_SEQ
    for m, c in zip(merged, clips):
        n = 8 * (W_TOTAL // N_SEQ) + 2                              # optimizer.py:425-437
        for key in ("final_estimated_seq", "mid_estimated_seq", "final_optimized_seq", "final_gt_seq", "mid_local_pose_seq"):
            assert tuple(m[key].shape) == (n, 15, 3), key
        est = m["final_estimated_seq"].cpu().numpy()
        x = c["estimated_local_skeleton"][:n]
        cam = c["camera_pose_list"][:n]
        direct = np.einsum("fab,fjb->fja", cam[:, :3, :3], x) + cam[:, None, :3, 3]
        assert np.abs(est - direct).max() < 1e-9
        assert np.array_equal(m["final_gt_seq"].cpu().numpy(), c["gt_global_skeleton"][:n])
        assert torch.isfinite(m["final_optimized_seq"]).all()


def test_sweep_point_of_1e4_windows_reproduces_the_1870_window_batch(full, vae_weights, camera):
    """BASELINE configs[4] (`bench.py --windows 10000` tiles the workload the same way): every window solved six times
    from the same noise in ONE batch of 11 220 windows — replicas share the clips' heat maps in HBM — gives, replica by
    replica, the rows of the 1870-window solve bit for bit."""
    from globalegomocap_b200.engine import Engine
    from globalegomocap_b200.pipeline import WindowBatch
    from globalegomocap_b200.vae_prep import PreparedVae
    R = 6
    eng = Engine(max_windows=W_TOTAL * R, max_history=24)
    try:
        prep = (PreparedVae(vae_weights[0], eng.device), PreparedVae(vae_weights[1], eng.device))
        kw = dict(full["kw"], local_vae_path=prep[0], global_vae_path=prep[1], engine=eng, outputs="optimized")
        batch = WindowBatch(eng, full["clips"]).replicate(R)
        assert batch.W == W_TOTAL * R == 11220
        out = full["gem"].solve_clips(batch, eps=full["eps"].repeat(R, 1, 1), **kw)
        torch.cuda.synchronize()
        for a, b in zip(_rows(out["sol"]), _rows(full["out"]["sol"])):
            for r in range(R):
                assert torch.equal(a[r * W_TOTAL:(r + 1) * W_TOTAL].cpu(), b.cpu()), r
        assert len(out["merged"]) == N_SEQ * R
    finally:
        eng.close()
