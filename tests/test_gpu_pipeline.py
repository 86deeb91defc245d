"""The drop-in entry points (BodyPoseOptimizer, main, CLI) on the GPU against the reference's
end-to-end golden run.  Needs a B200: `pytest -m gpu`."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from globalegomocap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, clip58, vae_weights):
    """Scratch tree laid out like a reference checkout: checkpoints under networks/logs/...,
    one clip under data/synth/clip0/test_data.pkl."""
    from globalegomocap_b200 import optimizer as gem
    root = tmp_path_factory.mktemp("gem_work")
    syn.save_checkpoint(vae_weights[0], str(root / gem.LOCAL_VAE_PATH))
    syn.save_checkpoint(vae_weights[1], str(root / gem.GLOBAL_VAE_PATH))
    syn.write_clip_pickle(clip58, str(root / "data" / "synth" / "clip0"))
    return root


def test_main_matches_reference_end_to_end(workdir, golden_dir, monkeypatch):
    from globalegomocap_b200 import optimizer as gem
    g = np.load(os.path.join(golden_dir, "main_mi3.npz"))
    monkeypatch.chdir(workdir)
    errors, est, mid_local, opt, gt = gem.main(
        "data/synth/clip0", camera_model_path=syn.DEFAULT_CAMERA_JSON, vae_weight=0.0, gmm_weight=0.0,
        smoothness_weight=0.001, bone_length_weight=0.01, weight_3d=0.01, reproj_weight=0.01, visualization=False,
        save=False, merge=True, final_smooth=True, max_iter=int(g["max_iter"]), eps=g["eps"])
    assert isinstance(est, list) and isinstance(mid_local, list) and isinstance(opt, np.ndarray)
    assert np.asarray(est).shape == np.asarray(mid_local).shape == opt.shape == np.asarray(gt).shape == (58, 15, 3)
    assert mid_local[0].dtype == np.float32 and est[0].dtype == np.float64
    np.testing.assert_allclose(np.asarray(est), g["final_estimated_seq"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.asarray(gt), g["final_gt_seq"], rtol=0, atol=1e-12)
    mid_mm = np.abs(np.asarray(mid_local) - g["mid_local_pose_seq"]).max(axis=(1, 2)) * 1000
    opt_mm = np.abs(opt - g["final_optimized_seq"]).max(axis=(1, 2)) * 1000
    print("mid_local mm per frame:", np.round(mid_mm, 3))
    print("optimized mm per frame:", np.round(opt_mm, 3))
    # the first and last windows' frames are bit-stable in the reference; windows whose line search
    # is ill-conditioned may part ways (tests/test_oracle_lbfgs.py explains and measures it)
    # (which window parts ways depends on the last bits of the closure: it differs between gemm modes)
    assert np.median(mid_mm) < 0.5 and (mid_mm < 0.5).mean() >= 0.5 and mid_mm.max() < 60
    assert np.median(opt_mm) < 0.5 and opt_mm.max() < 60
    assert len(errors) == 18
    for k in ("original_global_mpjpe", "original_camera_pos_error", "aligned_original_mpjpe",
              "bone_length_aligned_original_mpjpe"):
        np.testing.assert_allclose(errors[k], g["err__" + k], rtol=1e-9)


def test_body_pose_optimizer_single_window_api(workdir, golden_dir, clip58, monkeypatch):
    """The reference's per-window call: same signature, returns (10,15,3) float32."""
    from globalegomocap_b200 import optimizer as gem
    monkeypatch.chdir(workdir)
    g = np.load(os.path.join(golden_dir, "traces.npz"))
    est = clip58["estimated_local_skeleton"]
    opt = gem.BodyPoseOptimizer(camera_model_path=syn.DEFAULT_CAMERA_JSON, mean_skeleton=torch.from_numpy(est).float(),
                                vae_path=gem.LOCAL_VAE_PATH, latent_dim=2048, network_seq_len=10, seq_len=10,
                                windows_size=1, overlap_size=2, lr=2, max_iter=1)
    opt.set_weights(vae_weight=0.0, gmm_weight=0.0, smooth_weight=0.001 / 100, bone_length_weight=0.01,
                    weight_3d=0.01 / 10000, reproj_weight=0.01)
    s = int(g["starts"][1])
    res = opt.optimize_pose_seq_pytorch_LBFGS(est[s:s + 10], clip58["heatmap_list"][s:s + 10], est[s:s + 10].copy(),
                                              eps=g["eps"][1, 0])
    assert res.shape == (10, 15, 3) and res.dtype == np.float32
    assert np.abs(res - g["mi1_w1_local_pose"]).max() * 1000 < 0.05
    # total_loss(z) at the reference's recorded points
    for k in range(3):
        e = float(opt.total_loss(torch.from_numpy(g["mi2_w1_local_z"][k])))
        assert abs(e - g["mi2_w1_local_E"][k]) <= 1e-4 * abs(g["mi2_w1_local_E"][k])


def test_cli_runs_and_prints_reference_summary(workdir):
    env = dict(os.environ, PYTHONPATH=REPO)
    out = subprocess.run([sys.executable, os.path.join(REPO, "optimize_whole_sequence.py"), "--data_path",
                          "data/synth", "--camera", syn.DEFAULT_CAMERA_JSON, "--max_iter", "2", "--seed", "0"],
                         cwd=workdir, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "running data: data/synth/clip0" in out.stdout
    assert "Average optimized global pose mpjpe:" in out.stdout
    assert "joints error is:" in out.stdout


def test_many_clips_batch_equals_one_by_one(workdir, clip58, vae_weights, camera):
    """Clips optimised together (one launch for all their windows) give bit-identical results to
    clips optimised alone: windows never interact, clip boundaries are respected by stitching."""
    from globalegomocap_b200.engine import Engine
    from globalegomocap_b200.pipeline import SequenceOptimizer
    eng = Engine(max_windows=32)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    eng.set_vae(1, vae_weights[1])
    clip_b = syn.make_clip(30, seed=9)
    clip_c = {k: v[:17] for k, v in clip58.items()}          # 17 frames -> a single window, 7 frames dropped
    so = SequenceOptimizer(eng, max_iter=4)
    rng = np.random.default_rng(5)
    eps = rng.standard_normal((7 + 3 + 1, 2, 2048)).astype(np.float32)
    _, _, merged = so.run([clip58, clip_b, clip_c], eps=eps)
    assert [m["final_optimized_seq"].shape[0] for m in merged] == [58, 26, 10]
    for clip, sl, m in ((clip58, slice(0, 7), merged[0]), (clip_b, slice(7, 10), merged[1]),
                        (clip_c, slice(10, 11), merged[2])):
        _, _, alone = so.run([clip], eps=eps[sl])
        for k in ("final_optimized_seq", "mid_local_pose_seq", "final_estimated_seq"):
            assert torch.equal(alone[0][k], m[k]), k
    eng.close()


def test_slices_streams_graphs_and_pipelined_upload_do_not_change_results(vae_weights, camera):
    """The fused local -> transform -> global call gives bit-identical results whether the windows run as one
    slice or several concurrent ones, with or without CUDA-graph replay, through separate stage calls, and
    with the clips uploaded on a copy stream while the first slices already run, or with the heat maps left in
    pinned host memory and read through the energy kernel's texel cache."""
    from globalegomocap_b200.engine import Engine
    from globalegomocap_b200.pipeline import SequenceOptimizer, WindowBatch
    clips = [syn.make_clip(800, seed=40 + i) for i in range(3)]            # 99 windows each: one slice per clip
    pinned = [{k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in c.items()} for c in clips]
    eng = Engine(max_windows=300)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    eng.set_vae(1, vae_weights[1])
    so = SequenceOptimizer(eng, max_iter=5)
    W = 3 * 99
    eps = torch.randn(W, 2, 2048, generator=torch.Generator().manual_seed(3))

    def run(batch):
        sol = so.solve(batch, eps=eps)
        torch.cuda.synchronize()
        return sol["local"]["pose"].clone(), sol["glob"]["pose"].clone(), sol["glob"]["func_evals"].clone()

    eng.set_chunks(1)
    ref = run(WindowBatch(eng, clips))
    assert int(ref[2].min()) >= 2
    eng.set_chunks(3)
    multi = run(WindowBatch(eng, clips))
    piped = run(WindowBatch(eng, pinned, copy_stream=torch.cuda.Stream()))
    eng.set_slices(None)
    # heat maps left in pinned host memory, read over PCIe through the energy kernel's texel cache
    host_heat = torch.cat([c["heatmap_list"] for c in pinned]).pin_memory()
    eng.texel_cache_stats(True)
    zero_copy = run(WindowBatch(eng, pinned, host_heat=host_heat))
    lookups, fetched = eng.texel_cache_stats(False)
    assert lookups > 0 and 0 < fetched < 2 * lookups, (lookups, fetched)      # without the cache: 4 per lookup
    eng.set_texel_cache(1)                       # the cache on device-resident maps
    cached = run(WindowBatch(eng, clips))
    eng.set_texel_cache(-1)
    eng.set_chunks(1)
    batch = WindowBatch(eng, clips)                                         # separate stage calls (trace path)
    sol = so.solve(batch, eps=eps, want_trace=True)
    torch.cuda.synchronize()
    staged = (sol["local"]["pose"], sol["glob"]["pose"], sol["glob"]["func_evals"])
    for name, other in (("3 slices", multi), ("pipelined upload", piped), ("separate stages", staged),
                        ("zero-copy heat maps", zero_copy), ("texel cache", cached)):
        for a, b in zip(ref, other):
            assert torch.equal(a, b), name
    eng.close()


def test_main_batch_equals_per_clip_main(workdir, monkeypatch):
    """`main_batch` (every window of every clip of a dataset in ONE batched solve) returns, clip by
    clip, exactly what consecutive `main` calls return: windows are independent, the mean bone lengths are per
    clip, and one (sum W, 2, latent) noise draw consumes the torch generator like the per-clip draws do."""
    from globalegomocap_b200 import optimizer as gem
    monkeypatch.chdir(workdir)
    names = []
    for i, frames in enumerate((42, 26, 34)):
        d = os.path.join("data", "batch", "clip%d" % i)
        syn.write_clip_pickle(syn.make_clip(frames, seed=40 + i), str(workdir / d))
        names.append(d)
    kw = dict(camera_model_path=syn.DEFAULT_CAMERA_JSON, vae_weight=0.0, gmm_weight=0.0, smoothness_weight=0.001,
              bone_length_weight=0.01, weight_3d=0.01, reproj_weight=0.01, visualization=False, save=False, merge=True,
              final_smooth=True, max_iter=3)
    torch.manual_seed(5)
    single = [gem.main(d, **kw) for d in names]
    torch.manual_seed(5)
    batched = gem.main_batch(names, **kw)
    assert len(batched) == 3
    for (e1, est1, mid1, opt1, gt1), (e2, est2, mid2, opt2, gt2) in zip(single, batched):
        assert np.array_equal(np.asarray(mid1), np.asarray(mid2)) and np.array_equal(opt1, opt2)
        assert np.array_equal(np.asarray(est1), np.asarray(est2)) and np.array_equal(np.asarray(gt1), np.asarray(gt2))
        for k in e1:
            np.testing.assert_array_equal(e1[k], e2[k])


def test_main_batch_ingest_modes_are_bit_identical(workdir, monkeypatch, vae_weights):
    """`main_batch` lands the pickles in pinned memory once (load_clips) and solves with the heat maps
    read in place over PCIe (ingest 'zero_copy', the default), or uploaded piecewise on a copy stream ('upload');
    both give exactly what `solve_clips` gives on clips already resident in HBM, and `outputs='optimized'` is the
    same optimised sequence."""
    from globalegomocap_b200 import optimizer as gem
    from globalegomocap_b200.pipeline import WindowBatch
    monkeypatch.chdir(workdir)
    names = []
    for i, frames in enumerate((810, 26, 402)):              # 101 + 3 + 50 windows: several slices, ragged clips
        d = os.path.join("data", "ingest", "clip%d" % i)
        syn.write_clip_pickle(syn.make_clip(frames, seed=60 + i), str(workdir / d))
        names.append(d)
    kw = dict(camera_model_path=syn.DEFAULT_CAMERA_JSON, vae_weight=0.0, gmm_weight=0.0, smoothness_weight=0.001,
              bone_length_weight=0.01, weight_3d=0.01, reproj_weight=0.01, final_smooth=True, max_iter=4)
    W = 101 + 3 + 50
    eps = torch.randn(W, 2, 2048, generator=torch.Generator().manual_seed(8))
    zero = gem.main_batch(names, eps=eps, ingest="zero_copy", **kw)
    auto = gem.main_batch(names, eps=eps, **kw)
    up = gem.main_batch(names, eps=eps, ingest="upload", **kw)
    clips = gem.load_clips(names)
    assert clips.heat_all.is_pinned()
    eng = gem.shared_engine(W, 3)
    assert clips.planar == 2                                 # load_clips tiles the maps ([frames, J, H/4, W/8, 4, 8]) by default
    with pytest.raises(ValueError):                          # tiled maps declared planar
        WindowBatch(eng, [{k: v.cuda() for k, v in c.items()} for c in clips], planar=True)
    resident = WindowBatch(eng, [{k: v.cuda() for k, v in c.items()} for c in clips], planar=clips.planar)
    res = gem.solve_clips(resident, eps=eps, **kw)
    hwc = gem.load_clips(names, planar=False)                # the pickle's own layout gives the same bits
    res_hwc = gem.solve_clips(WindowBatch(eng, [{k: v.cuda() for k, v in c.items()} for c in hwc]), eps=eps, **kw)
    for a, b in zip(res["merged"], res_hwc["merged"]):
        assert torch.equal(a["final_optimized_seq"], b["final_optimized_seq"])
    opt_only = gem.solve_clips(resident, eps=eps, outputs="optimized", **kw)
    torch.cuda.synchronize()
    for i in range(3):
        want = res["merged"][i]
        for other in (zero, auto, up):
            errors, est, mid, opt, gt = other[i]
            assert np.array_equal(opt, want["final_optimized_seq"].cpu().numpy())
            assert np.array_equal(np.asarray(mid), want["mid_local_pose_seq"].cpu().numpy())
            assert np.array_equal(np.asarray(est), want["final_estimated_seq"].cpu().numpy())
        assert torch.equal(opt_only["merged"][i]["final_optimized_seq"], want["final_optimized_seq"])
    with pytest.raises(NotImplementedError):                 # rejected before any work is done
        gem.main_batch(names, eps=eps, save=True, **kw)


def test_replicated_windows_share_inputs(vae_weights, camera):
    """WindowBatch.replicate (the 1e5-window sweep point of BASELINE configs[4]): replicas fed the same noise give
    the same result as the original windows; each replica is stitched as its own clip."""
    from globalegomocap_b200 import optimizer as gem
    from globalegomocap_b200.engine import Engine
    from globalegomocap_b200.pipeline import WindowBatch
    from globalegomocap_b200.vae_prep import PreparedVae
    clips = [syn.make_clip(50, seed=70), syn.make_clip(26, seed=71)]
    eng = Engine(max_windows=64)
    prep = (PreparedVae(vae_weights[0], eng.device), PreparedVae(vae_weights[1], eng.device))
    kw = dict(camera_model_path=camera, final_smooth=True, max_iter=3, local_vae_path=prep[0], global_vae_path=prep[1],
              engine=eng, outputs="optimized")
    eps = torch.randn(9, 2, 2048, generator=torch.Generator().manual_seed(2))
    one = gem.solve_clips(WindowBatch(eng, clips), eps=eps, **kw)
    rep = gem.solve_clips(WindowBatch(eng, clips).replicate(3), eps=eps.repeat(3, 1, 1), **kw)
    assert len(rep["merged"]) == 6 and rep["batch"].W == 27
    for r in range(3):
        for i in range(2):
            assert torch.equal(rep["merged"][2 * r + i]["final_optimized_seq"], one["merged"][i]["final_optimized_seq"])
    eng.close()


def test_main_on_the_1k_frame_sequence_matches_the_reference_run(workdir, golden_dir, monkeypatch):
    """BASELINE configs[0]: `optimizer.main` on one synthetic 1000-frame sequence (124 windows, both stages, the
    reference's own max_iter = 25) against the UNMODIFIED reference's run of the same call
    (tests/golden/make_golden_main1k.py: 140 s on eight CPU threads; the inputs are regenerated from their seeds).
    Bit-level agreement of a free-running 25-iteration solve is statistical for the reference itself (DESIGN.md section 2,
    tests/test_gpu_dist.py) — with random-init VAEs the energy is nearly flat along most of the latent and the 25
    iterations end tens of millimetres apart even between two runs of the reference — so the golden also holds the
    reference's run at ONE thread, and the yardstick is the reference's own spread: the un-optimised sequences and
    their metrics agree to round-off, the optimised sequences deviate per frame like the reference deviates from
    itself, and the error metrics the reference's CLI prints (optimize_whole_sequence.py:90-117) agree to 0.1 %."""
    from globalegomocap_b200 import optimizer as gem
    g = np.load(os.path.join(golden_dir, "main_1k_mi25.npz"))
    n, W = int(g["n_frames"]), int(g["windows"])
    clip = syn.make_clip(n, seed=int(g["seed"]))
    eps = np.random.default_rng(int(g["eps_seed"])).standard_normal((W, 2, 2048)).astype(np.float32)
    syn.write_clip_pickle(clip, str(workdir / "data" / "synth" / "clip1k"))
    monkeypatch.chdir(workdir)
    errors, est, mid_local, opt, gt = gem.main(
        "data/synth/clip1k", camera_model_path=syn.DEFAULT_CAMERA_JSON, vae_weight=0.0, gmm_weight=0.0,
        smoothness_weight=0.001, bone_length_weight=0.01, weight_3d=0.01, reproj_weight=0.01, visualization=False,
        save=False, merge=True, final_smooth=True, max_iter=25, eps=eps)
    frames = 8 * W + 2
    assert np.asarray(est).shape == np.asarray(mid_local).shape == opt.shape == np.asarray(gt).shape == (frames, 15, 3)
    np.testing.assert_allclose(np.asarray(est), g["final_estimated_seq"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.asarray(gt), g["final_gt_seq"], rtol=0, atol=1e-12)
    for k in ("original_global_mpjpe", "original_camera_pos_error", "original_aligned_camera_pos_error",
              "original_aligned_global_mpjpe", "aligned_original_mpjpe", "bone_length_aligned_original_mpjpe"):
        np.testing.assert_allclose(errors[k], g["err__" + k], rtol=1e-9)
    q = (25, 50, 75, 90, 99, 100)

    def mm(a, b):
        return np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max(axis=(1, 2)) * 1000

    # per-frame deviation from the reference's 8-thread run: this library, and the reference's own 1-thread run
    mid_mm, mid_self = mm(mid_local, g["mid_local_pose_seq"]), mm(g["mid_local_pose_seq_one_thread"], g["mid_local_pose_seq"])
    opt_mm, opt_self = mm(opt, g["final_optimized_seq"]), mm(g["final_optimized_seq_one_thread"], g["final_optimized_seq"])
    print("percentiles", q)
    print("local-stage result, mm per frame:  CUDA", np.round(np.percentile(mid_mm, q), 3),
          " reference 1 vs 8 threads", np.round(np.percentile(mid_self, q), 3))
    print("optimised sequence, mm per frame:  CUDA", np.round(np.percentile(opt_mm, q), 3),
          " reference 1 vs 8 threads", np.round(np.percentile(opt_self, q), 3))
    rel, rel_self = {}, {}
    for k in errors:
        if "original" in k:
            continue
        ref = float(np.mean(g["err__" + k]))
        rel[k] = float(np.mean(errors[k])) / ref - 1.0
        rel_self[k] = float(np.mean(g["err1__" + k])) / ref - 1.0
    print("optimised metrics, CUDA / reference - 1:", {k: round(v, 5) for k, v in rel.items()})
    print("optimised metrics, reference 1 thread / 8 threads - 1:", {k: round(v, 5) for k, v in rel_self.items()})
    worst, worst_self = max(abs(v) for v in rel.values()), max(abs(v) for v in rel_self.values())
    print("worst metric deviation: CUDA %.5f, reference self-noise %.5f" % (worst, worst_self))
    # measured (B200, default arithmetic): optimised sequence median 26.0 mm / p90 45.3 mm against the reference's own
    # 24.4 / 40.7 mm between one and eight threads; local stage p90 187 mm against 173 mm; worst printed metric 0.109 %
    # against 0.128 %
    assert np.median(opt_mm) < 1.5 * np.median(opt_self) and np.percentile(opt_mm, 90) < 1.5 * np.percentile(opt_self, 90)
    assert np.percentile(mid_mm, 90) < 1.5 * np.percentile(mid_self, 90) and mid_mm.max() < 2.0 * mid_self.max()
    assert worst < 3.0 * worst_self


def test_save_pose_writes_the_references_result_pickle(workdir, monkeypatch):
    """`save_pose=True` (optimizer.py:469-483): out/<dataset>/<sequence>/result_pose.pkl with the reference's four keys
    and container types — lists of per-frame (15, 3) arrays, the optimised sequence an ndarray when `final_smooth` — holding
    exactly the sequences `main` returns.  `save=True` / `visualization=True` (open3d meshes) are refused before any work."""
    import pickle

    from globalegomocap_b200 import optimizer as gem
    monkeypatch.chdir(workdir)
    kw = dict(camera_model_path=syn.DEFAULT_CAMERA_JSON, vae_weight=0.0, gmm_weight=0.0, smoothness_weight=0.001,
              bone_length_weight=0.01, weight_3d=0.01, reproj_weight=0.01, max_iter=2)
    for smooth in (True, False):
        torch.manual_seed(3)
        errors, est, mid_local, opt, gt = gem.main("data/synth/clip0", final_smooth=smooth, save_pose=True, **kw)
        path = workdir / "out" / "synth" / "clip0" / "result_pose.pkl"
        with open(path, "rb") as f:
            saved = pickle.load(f)
        assert sorted(saved) == ["estimated_pose", "gt_pose", "mid_optimized_pose", "optimized_pose"]
        assert isinstance(saved["estimated_pose"], list) and isinstance(saved["gt_pose"], list)
        assert isinstance(saved["mid_optimized_pose"], list) and saved["mid_optimized_pose"][0].shape == (15, 3)
        assert isinstance(saved["optimized_pose"], np.ndarray if smooth else list)
        assert np.array_equal(np.asarray(saved["estimated_pose"]), np.asarray(est))
        assert np.array_equal(np.asarray(saved["optimized_pose"]), np.asarray(opt))
        assert np.array_equal(np.asarray(saved["gt_pose"]), np.asarray(gt))
        assert np.asarray(saved["mid_optimized_pose"]).shape == np.asarray(est).shape
        os.remove(path)
    for bad in (dict(save=True), dict(visualization=True)):
        with pytest.raises(NotImplementedError):
            gem.main("data/synth/clip0", **bad, **kw)
    assert not (workdir / "out" / "synth" / "clip0" / "result_pose.pkl").exists()


def test_per_term_methods_of_the_optimizer_class_match_reference(workdir, golden_dir, clip58, monkeypatch):
    """`pose_energy_3d`, `smooth_accelerate`, `bone_length_energy`, `vae_energy`,
    `reprojection_energy_heatmap_fast` and `calculate_bone_length` keep the reference's names and arguments
    (optimizer.py:89-94, 139-149, 172-177, 202-218) and give the reference's values (golden per-term energies of 15
    cases, 5e-6) at the state an optimize call leaves; their weighted sum is `total_loss`'s combination
    (optimizer.py:239-240)."""
    from globalegomocap_b200 import optimizer as gem
    monkeypatch.chdir(workdir)
    g = np.load(os.path.join(golden_dir, "energy.npz"))
    opt = gem.BodyPoseOptimizer(camera_model_path=syn.DEFAULT_CAMERA_JSON,
                                mean_skeleton=torch.from_numpy(clip58["estimated_local_skeleton"]).float(),
                                vae_path=gem.LOCAL_VAE_PATH, latent_dim=2048, network_seq_len=10, seq_len=10,
                                windows_size=1, overlap_size=2, lr=2, max_iter=2)
    np.testing.assert_allclose(opt.mean_bone_length.cpu().numpy(), g["mean_bone_length"], rtol=1e-6)
    bl = opt.calculate_bone_length(torch.from_numpy(clip58["estimated_local_skeleton"][:10]).float().view(10, 45))
    assert tuple(bl.shape) == (10, 15) and float(bl[:, 0].abs().max()) == 0.0
    names = [str(n) for n in g["names"] if str(n).startswith("all__")]
    w = [float(v) for v in g["all__weights"]]                       # {w3d, smooth, bone, vae, reproj}
    opt.set_weights(vae_weight=w[3], gmm_weight=0.0, smooth_weight=w[1], bone_length_weight=w[2], weight_3d=w[0],
                    reproj_weight=w[4])
    checked = 0
    for n in names:
        if n.endswith("edges") or n.endswith("dense_near"):
            continue                                                # (cases on synthetic dense maps: test_gpu_kernels.py)
        s = int(g[f"{n}__start"])
        x = torch.from_numpy(g[f"{n}__x"]).float()
        # the state optimize_pose_seq_pytorch_LBFGS leaves (optimizer.py:247-252)
        opt.initial_pose = torch.from_numpy(clip58["estimated_local_skeleton"][s:s + 10]).float()
        opt.heatmap_seq = torch.from_numpy(clip58["heatmap_list"][s:s + 10]).float()
        got = {"e3d": opt.pose_energy_3d(x), "smooth": opt.smooth_accelerate(x), "bone": opt.bone_length_energy(x),
               "vae": opt.vae_energy(x), "reproj": opt.reprojection_energy_heatmap_fast(x)}
        for t, v in got.items():
            ref = float(g[f"{n}__E_{t}"])
            assert abs(float(v) - ref) <= 5e-6 * max(abs(ref), 1.0), (n, t, float(v), ref)
        total = (w[0] * float(got["e3d"]) + w[1] * float(got["smooth"]) + w[2] * float(got["bone"]) + w[3] * float(got["vae"])
                 + w[4] * float(got["reproj"]))
        ref = float(g[f"{n}__E_total"])
        assert abs(total - ref) <= 2e-5 * max(abs(ref), 1e-2), (n, total, ref)
        checked += 1
    assert checked >= 3
    fresh = gem.BodyPoseOptimizer(camera_model_path=syn.DEFAULT_CAMERA_JSON,
                                  mean_skeleton=torch.from_numpy(clip58["estimated_local_skeleton"]).float(),
                                  vae_path=gem.LOCAL_VAE_PATH, latent_dim=2048, network_seq_len=10, seq_len=10, engine=opt.engine)
    with pytest.raises(gem.GemError):
        fresh.pose_energy_3d(x)                                     # no optimize call yet: no initial pose
    assert float(fresh.vae_energy(torch.ones(1, 2048))) == 2048.0   # the one-liner on a latent-shaped argument


def test_cli_summary_matches_the_references_cli(workdir):
    """The reference's `optimize_whole_sequence.py` run UNMODIFIED on a three-clip dataset (tests/golden/make_golden_cli.py,
    `torch.manual_seed(0)` set before it) against this repository's CLI with `--seed 0` on the same dataset, batched and
    clip by clip: the same lines in the same order (clips in natural order), the un-optimised figures to round-off, the
    optimised ones — 25 free-running iterations on 9 windows, statistical for the reference itself — within 3 %."""
    import json
    import re
    with open(os.path.join(REPO, "tests", "golden", "cli_summary.json")) as f:
        g = json.load(f)
    root = workdir / "data" / "synth_cli"
    for name, frames, seed in g["clips"]:
        syn.write_clip_pickle(syn.make_clip(int(frames), seed=int(seed)), str(root / name))
    env = dict(os.environ, PYTHONPATH=REPO)
    for batch in ("True", "False"):
        out = subprocess.run([sys.executable, os.path.join(REPO, "optimize_whole_sequence.py"), "--data_path", "data/synth_cli",
                              "--camera", syn.DEFAULT_CAMERA_JSON, "--seed", str(g["seed"]), "--batch_clips", batch],
                             cwd=workdir, env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith(("Average", "running data", "-----"))]
        assert [ln.split(":")[0] for ln in lines] == [ln.split(":")[0] for ln in g["lines"]]
        assert [ln for ln in lines if ln.startswith("running")] == [ln for ln in g["lines"] if ln.startswith("running")]
        worst = 0.0
        for ln in lines:
            m = re.match(r"(Average [^:]+): (.*)$", ln)
            if not m:
                continue
            ours, ref = float(m.group(2)), g["summary"][m.group(1)]
            if "original" in m.group(1):
                assert abs(ours - ref) <= 1e-9 * abs(ref), (ln, ref)
            else:
                worst = max(worst, abs(ours / ref - 1.0))
        je = re.search(r"joints error is: \[([^\]]*)\]", out.stdout)
        ours_je = np.array([float(t) for t in je.group(1).split()])
        assert ours_je.shape == (15,)
        worst_je = float(np.abs(ours_je / np.array(g["joints_error"]) - 1.0).max())
        print("batch_clips", batch, "worst optimised summary deviation %.4f, worst per-joint error deviation %.4f" % (worst, worst_je))
        assert worst < 0.03 and worst_je < 0.10
