"""Per-ctx state on the B200: calibration (11 and 14 coefficients), swapped VAE weights under cached CUDA graphs,
fp16 range status, input validation.  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch

from globalegomocap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

CAM14_JSON = syn.DEFAULT_CAMERA_JSON.replace("calibration.json", "calibration_new.json")
W_LOCAL = (0.01 / 10000, 0.001 / 100, 0.01, 0.0, 0.01)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_energy_with_the_14_coefficient_calibration_matches_reference(golden_dir, clip58):
    """pose_fisheye_fisheye.calibration_new.json (FishEyeCalibrated.py:8-14): the kernel's polynomial loop runs to
    n_poly = 14; reprojection term, weighted total and both gradients vs the unmodified reference's autograd."""
    from globalegomocap_b200.engine import Engine, energy_weights
    g = np.load(os.path.join(golden_dir, "energy_cam14.npz"))
    eng = Engine(max_windows=16)
    eng.set_camera_json(CAM14_JSON)
    assert int(g["n_poly"]) == 14
    heat1 = syn.dense_heat_window(2)
    mb = g["mean_bone_length"]
    names = [str(n) for n in g["names"]]
    for wname in ("local", "all"):
        sub = [n for n in names if n.startswith(wname + "__")]
        xs = np.stack([g[f"{n}__x"] for n in sub])
        starts = [int(g[f"{n}__start"]) for n in sub]
        x0 = np.stack([clip58["estimated_local_skeleton"][s:s + 10] for s in starts]).astype(np.float32)
        heat = np.concatenate([heat1] * len(sub))
        fb = np.arange(len(sub), dtype=np.int64) * 10
        w = g[f"{wname}__weights"]
        E, terms, grad, status = eng.energy_grad(xs, x0, heat, fb, np.zeros(len(sub), np.int32), mb, energy_weights(*w))
        E, terms, grad = E.cpu().numpy(), terms.cpu().numpy(), grad.cpu().numpy()
        assert int(status.sum()) == 0
        for i, n in enumerate(sub):
            e_ref = float(g[f"{n}__E_reproj"])
            assert abs(terms[i, 4] - e_ref) <= 5e-6 * max(abs(e_ref), 1.0), (n, terms[i, 4], e_ref)
            scale = float(np.abs(w * terms[i].astype(np.float64)).sum())           # the total is a cancelling sum
            assert abs(E[i] - float(g[f"{n}__E_total"])) <= 2e-5 * max(scale, 1e-2), (n, E[i])
            assert _rel(grad[i], g[f"{n}__G_total"]) < 5e-5, (n, _rel(grad[i], g[f"{n}__G_total"]))
        onehot = energy_weights(0, 0, 0, 0, 1.0)
        _, _, grad, _ = eng.energy_grad(xs, x0, heat, fb, np.zeros(len(sub), np.int32), mb, onehot)
        for i, n in enumerate(sub):
            assert _rel(grad[i].cpu().numpy(), g[f"{n}__G_reproj"]) < 5e-5, n
    eng.close()


def test_two_ctxs_keep_their_own_calibration(clip58):
    """camera and skeleton used to live in process-global __constant__ memory, so the last
    gem_ctx_set_camera won for every ctx.  They are per-ctx kernel parameters now."""
    from globalegomocap_b200.engine import Engine, energy_weights
    a, b = Engine(max_windows=4), Engine(max_windows=4)
    a.set_camera_json(syn.DEFAULT_CAMERA_JSON)
    x = clip58["estimated_local_skeleton"][:10][None].astype(np.float32)
    heat = syn.dense_heat_window(3)
    args = (x, x, heat, np.zeros(1, np.int64), np.zeros(1, np.int32), np.full(15, 0.25, np.float32),
            energy_weights(0, 0, 0, 0, 1.0))
    e_a0 = float(a.energy_grad(*args)[0][0])
    b.set_camera_json(CAM14_JSON)                      # must not leak into ctx a
    e_b = float(b.energy_grad(*args)[0][0])
    e_a1 = float(a.energy_grad(*args)[0][0])
    assert e_a0 == e_a1 and e_a0 != e_b
    a.close(), b.close()


def test_swapping_vae_weights_invalidates_cached_graphs(clip58, camera):
    """the CUDA graph of rounds 1.. is cached per configuration; gem_ctx_set_vae must drop it, or the
    replayed rounds keep reading the previous weights (and their freed prepared slabs)."""
    from globalegomocap_b200.engine import Engine, energy_weights, lbfgs_params
    from oracle import energy_np as en
    bias = syn.mean_pose_bias(clip58)
    sd_a = syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias)
    sd_b = syn.make_vae_state_dict(12, perturb_bn=True, pose_bias=bias)
    est, heat = clip58["estimated_local_skeleton"], clip58["heatmap_list"]
    mb = en.mean_bone_length(est)
    starts = [0, 8, 16, 24]
    x0 = np.stack([est[s:s + 10] for s in starts]).astype(np.float32)
    eps = np.random.default_rng(5).standard_normal((4, 2048)).astype(np.float32)

    def solve(eng):
        r = eng.solve_stage(0, x0, heat, np.asarray(starts, np.int64), np.zeros(4, np.int32), mb, eps,
                            energy_weights(*W_LOCAL), lbfgs_params(max_iter=5), want_trace=True)
        return r["pose"].cpu().numpy(), r["trace"].cpu().numpy()

    fresh = Engine(max_windows=4)
    fresh.set_camera(*camera)
    fresh.set_vae(0, sd_b)
    want_pose, want_trace = solve(fresh)
    fresh.close()
    eng = Engine(max_windows=4)
    eng.set_camera(*camera)
    eng.set_vae(0, sd_a)
    solve(eng)                                          # captures the graph with sd_a's weights
    eng.set_vae(0, sd_b)
    got_pose, got_trace = solve(eng)
    assert np.array_equal(got_pose, want_pose) and np.array_equal(got_trace, want_trace, equal_nan=True)
    eng.close()


def test_fp16_range_overflow_raises_a_status_bit(clip58, camera):
    """The split-fp16 operands saturate at 65504: a decoder whose activations leave that range
    must raise GEM_WIN_F16_RANGE (the shim turns it into GemError), not degrade silently; the 3xTF32 mode has fp32's
    range and stays clean."""
    from globalegomocap_b200 import optimizer as gem
    from globalegomocap_b200.engine import Engine, GemError, energy_weights, lbfgs_params
    from oracle import energy_np as en
    bias = syn.mean_pose_bias(clip58)
    sd = syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias)
    sd["decoder_input.weight"] = sd["decoder_input.weight"] * np.float32(4e4)      # |T*256 activations| >> 65504
    est, heat = clip58["estimated_local_skeleton"], clip58["heatmap_list"]
    mb = en.mean_bone_length(est)
    x0 = est[:10][None].astype(np.float32)
    eps = np.random.default_rng(6).standard_normal((1, 2048)).astype(np.float32)
    eng = Engine(max_windows=4)
    eng.set_camera(*camera)
    eng.set_vae(0, sd)
    args = (0, x0, heat, np.zeros(1, np.int64), np.zeros(1, np.int32), mb, eps, energy_weights(*W_LOCAL),
            lbfgs_params(max_iter=1))
    res = eng.solve_stage(*args)
    assert int(res["status"][0]) & gem.GEM_WIN_F16_RANGE
    with pytest.raises(GemError, match="fp16"):
        gem._raise_on_status(res["status"])
    eng.set_gemm_mode(1)
    res = eng.solve_stage(*args)
    assert int(res["status"][0]) & gem.GEM_WIN_F16_RANGE == 0
    eng.close()


def test_heat_map_shape_is_validated(clip58, camera):
    """the C ABI takes a raw map pointer; the shim checks resolution, joint count and frame range."""
    from globalegomocap_b200.engine import Engine, GemError, energy_weights
    eng = Engine(max_windows=4)
    eng.set_camera(*camera)
    x = clip58["estimated_local_skeleton"][:10][None].astype(np.float32)
    w = energy_weights(0, 0, 0, 0, 1.0)
    mbone = np.full(15, 0.25, np.float32)
    with pytest.raises(GemError, match="heat"):
        eng.energy_grad(x, x, np.zeros((10, 32, 32, 15), np.float32), np.zeros(1, np.int64), np.zeros(1, np.int32), mbone, w)
    with pytest.raises(GemError, match="frame"):
        eng.energy_grad(x, x, np.zeros((10, 64, 64, 15), np.float32), np.asarray([5], np.int64), np.zeros(1, np.int32), mbone, w)
    eng.close()


def test_empty_and_too_short_inputs(tmp_path, clip58, vae_weights, camera, monkeypatch):
    """Edge sizes: a call with zero windows is a no-op that reports success (every entry point); more windows than
    the ctx was sized for is refused; a clip shorter than one window has no windows (`range(0, N - 10 + 1, 8)` is empty,
    optimizer.py:370) — `solve_clips` stitches nothing for it beside the other clips' results, `main` refuses it
    instead of failing inside the merge like the reference's empty `merge_batches` (optimizer.py:425-437)."""
    from globalegomocap_b200 import optimizer as gem
    from globalegomocap_b200.engine import Engine, GemError, energy_weights, lbfgs_params
    from globalegomocap_b200.vae_prep import PreparedVae
    eng = Engine(max_windows=8)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    eng.set_vae(1, vae_weights[1])
    heat = clip58["heatmap_list"][:10]
    mbone = np.full(15, 0.25, np.float32)
    w = energy_weights(*W_LOCAL)
    x0 = np.zeros((0, 10, 15, 3), np.float32)
    E, terms, grad, status = eng.energy_grad(x0, x0, heat, np.zeros(0, np.int64), np.zeros(0, np.int32), mbone, w)
    assert E.shape == (0,) and grad.shape == (0, 10, 15, 3) and status.shape == (0,)
    assert eng.decode(0, np.zeros((0, 2048), np.float32)).shape == (0, 10, 15, 3)
    r = eng.solve_stage(0, x0, heat, np.zeros(0, np.int64), np.zeros(0, np.int32), mbone, np.zeros((0, 2048), np.float32), w,
                        lbfgs_params(max_iter=3))
    torch.cuda.synchronize()
    assert r["pose"].shape == (0, 10, 15, 3)
    too_many = np.zeros((9, 10, 15, 3), np.float32)
    with pytest.raises(GemError, match="capacity"):
        eng.energy_grad(too_many, too_many, heat, np.zeros(9, np.int64), np.zeros(9, np.int32), mbone, w)
    # a 9-frame clip between two real ones
    short = {k: v[:9] for k, v in clip58.items()}
    one = {k: v[:10] for k, v in clip58.items()}
    prep = (PreparedVae(vae_weights[0], eng.device), PreparedVae(vae_weights[1], eng.device))
    eps = torch.randn(2, 2, 2048, generator=torch.Generator().manual_seed(1))
    out = gem.solve_clips([one, short, one], camera, max_iter=2, eps=eps, local_vae_path=prep[0], global_vae_path=prep[1],
                          engine=eng)
    assert out["merged"][1] is None and out["batch"].W == 2
    assert out["merged"][0]["final_optimized_seq"].shape == (10, 15, 3) == out["merged"][2]["final_optimized_seq"].shape
    syn.write_clip_pickle(short, str(tmp_path / "short"))
    with pytest.raises(ValueError, match="shorter than one window"):
        gem.main(str(tmp_path / "short"), camera_model_path=syn.DEFAULT_CAMERA_JSON, vae_weight=0.0, gmm_weight=0.0,
                 smoothness_weight=0.001, bone_length_weight=0.01, weight_3d=0.01, reproj_weight=0.01, engine=eng,
                 local_vae_path=prep[0], global_vae_path=prep[1])
    eng.close()
