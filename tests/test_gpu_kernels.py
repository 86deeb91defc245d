"""CUDA kernels (through the C ABI) against the reference goldens and the CPU oracle.
All tests here need a B200: run with `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch

from globalegomocap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

TERMS = ("e3d", "smooth", "bone", "vae", "reproj")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def engine(vae_weights, camera):
    from globalegomocap_b200.engine import Engine
    eng = Engine(max_windows=64)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    eng.set_vae(1, vae_weights[1])
    yield eng
    eng.close()


def _heat_for(name, clip58, start):
    if name.endswith("edges") or name.endswith("dense_near"):
        return syn.dense_heat_window(1)
    return clip58["heatmap_list"][start:start + 10]


def test_energy_terms_and_gradients_match_reference(engine, golden_dir, clip58):
    """Per-term energies, per-term gradients (weights one-hot) and the weighted total for the
    15 golden cases, all evaluated in ONE launch per weight set (W = cases)."""
    from globalegomocap_b200.engine import energy_weights
    g = np.load(os.path.join(golden_dir, "energy.npz"))
    mb = g["mean_bone_length"]
    names = [str(n) for n in g["names"]]
    for wname in ("local", "global", "all"):
        sub = [n for n in names if n.startswith(wname + "__")]
        xs = np.stack([g[f"{n}__x"] for n in sub])
        starts = [int(g[f"{n}__start"]) for n in sub]
        x0 = np.stack([clip58["estimated_local_skeleton"][s:s + 10] for s in starts]).astype(np.float32)
        heat = np.concatenate([_heat_for(n, clip58, s) for n, s in zip(sub, starts)])      # [cases*10,64,64,15]
        frame_base = np.arange(len(sub), dtype=np.int64) * 10
        clip_idx = np.zeros(len(sub), dtype=np.int32)
        w = g[f"{wname}__weights"]
        E, terms, grad, status = engine.energy_grad(xs, x0, heat, frame_base, clip_idx, mb, energy_weights(*w))
        E, terms, grad = E.cpu().numpy(), terms.cpu().numpy(), grad.cpu().numpy()
        assert int(status.sum()) == 0
        for i, n in enumerate(sub):
            for k, t in enumerate(TERMS):
                if t == "reproj" and w[4] == 0:
                    continue
                e_ref = float(g[f"{n}__E_{t}"])
                assert abs(terms[i, k] - e_ref) <= 5e-6 * max(abs(e_ref), 1.0), (n, t, terms[i, k], e_ref)
            e_ref = float(g[f"{n}__E_total"])
            assert abs(E[i] - e_ref) <= 2e-5 * max(abs(e_ref), 1e-2), (n, E[i], e_ref)
            assert _rel(grad[i], g[f"{n}__G_total"]) < 5e-5, (n, _rel(grad[i], g[f"{n}__G_total"]))
    # per-term gradients: one-hot weights isolate each term's analytic gradient
    sub = [n for n in names if n.startswith("all__")]
    xs = np.stack([g[f"{n}__x"] for n in sub])
    starts = [int(g[f"{n}__start"]) for n in sub]
    x0 = np.stack([clip58["estimated_local_skeleton"][s:s + 10] for s in starts]).astype(np.float32)
    heat = np.concatenate([_heat_for(n, clip58, s) for n, s in zip(sub, starts)])
    for k, t in enumerate(TERMS):
        onehot = [0.0] * 5
        onehot[k] = 1.0
        _, _, grad, _ = engine.energy_grad(xs, x0, heat, np.arange(len(sub), dtype=np.int64) * 10,
                                           np.zeros(len(sub), np.int32), mb, energy_weights(*onehot))
        grad = grad.cpu().numpy()
        for i, n in enumerate(sub):
            assert _rel(grad[i], g[f"{n}__G_{t}"]) < 5e-5, (n, t, _rel(grad[i], g[f"{n}__G_{t}"]))


def test_energy_norm_zero_sets_status(engine, clip58):
    from globalegomocap_b200.engine import energy_weights
    x = clip58["estimated_local_skeleton"][:20].reshape(2, 10, 15, 3).astype(np.float32).copy()
    x[1, 3, 7, :2] = 0.0
    heat = clip58["heatmap_list"][:20]
    _, _, _, status = engine.energy_grad(x, x, heat, np.array([0, 10]), np.zeros(2, np.int32),
                                         np.ones(15, np.float32), energy_weights(0, 0, 0, 0, 1.0))
    assert status.cpu().tolist() == [0, 1]


def test_energy_batch_independent_and_bulk_path_equal(engine, clip58):
    """The same window gives bit-identical results wherever it sits in the batch, for odd and
    even batch sizes (TMA bulk-copy path for window pairs vs the scalar tail path)."""
    from globalegomocap_b200.engine import energy_weights
    rng = np.random.default_rng(0)
    x0 = clip58["estimated_local_skeleton"][:10].astype(np.float32)
    x = x0 + 0.01 * rng.standard_normal(x0.shape).astype(np.float32)
    heat = clip58["heatmap_list"][:10]
    mb = np.full(15, 0.25, np.float32)
    w = energy_weights(0.01, 0.02, 0.05, 0.003, 0.04)
    outs = []
    for W in (1, 2, 5, 8):
        E, terms, grad, _ = engine.energy_grad(np.stack([x] * W), np.stack([x0] * W), heat, np.zeros(W, np.int64),
                                               np.zeros(W, np.int32), mb, w)
        E, grad = E.cpu().numpy(), grad.cpu().numpy()
        assert (E == E[0]).all() and (grad == grad[0]).all()
        outs.append((E[0], grad[0]))
    for e, gr in outs[1:]:
        assert e == outs[0][0] and (gr == outs[0][1]).all()


@pytest.mark.parametrize("layout", ["hwc", "planar", "tiled"])
def test_energy_compile_time_shape_matches_runtime_shape(engine, clip58, layout):
    """The energy kernel instantiated for the reference's configuration (10 x 15 joints, 64 x 64 maps, 11 polynomial
    coefficients as compile-time constants: multiplications for the divisions by J, shifts for the texel addresses, an
    unrolled polynomial) evaluates the same operations in the same order as the runtime-shape kernel: bit-identical
    energies, terms and gradients in every heat-map layout, with joints at the map border and outside it."""
    import ctypes as C
    from globalegomocap_b200.engine import energy_weights
    rng = np.random.default_rng(5)
    W = 7
    x0 = np.stack([clip58["estimated_local_skeleton"][8 * i:8 * i + 10] for i in range(W)]).astype(np.float32)
    x = x0 + 0.02 * rng.standard_normal(x0.shape).astype(np.float32)
    x[3] *= 1.8                                    # projections that leave the map
    heat = np.ascontiguousarray(clip58["heatmap_list"][:8 * W + 10])
    if layout == "planar":
        heat_in = np.ascontiguousarray(heat.transpose(0, 3, 1, 2))
    elif layout == "tiled":
        n = heat.shape[0]
        heat_in = np.ascontiguousarray(heat.transpose(0, 3, 1, 2).reshape(n, 15, 16, 4, 8, 8).transpose(0, 1, 2, 4, 3, 5))
    else:
        heat_in = heat
    fb = np.arange(W, dtype=np.int64) * 8
    mb = np.full(15, 0.25, np.float32)
    w = energy_weights(0.01, 0.02, 0.05, 0.003, 0.04)
    lib = engine.lib
    lib.gem_debug_energy_fixed.argtypes = [C.c_int]
    code = {"hwc": 0, "planar": 1, "tiled": 2}[layout]
    out = {}
    try:
        engine.set_heat_layout(code)
        for fixed in (0, 1):
            lib.gem_debug_energy_fixed(fixed)
            E, terms, grad, status = engine.energy_grad(x, x0, heat_in, fb, np.zeros(W, np.int32), mb, w)
            out[fixed] = (E.cpu().numpy(), terms.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy())
    finally:
        lib.gem_debug_energy_fixed(-1)
        engine.set_heat_layout(0)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert np.abs(out[1][2]).max() > 0


def test_vae_decode_vjp_encode_match_reference(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "vae.npz"))
    pose = engine.decode(0, g["z"]).cpu().numpy()
    assert pose.shape == (3, 10, 15, 3)
    assert _rel(pose, g["pose"]) < 2e-5
    dz = engine.decode_vjp(0, g["upstream"]).cpu().numpy()
    assert _rel(dz, g["dz"]) < 2e-4
    z0, mu, std = engine.encode(0, g["enc_in"], np.zeros((3, 2048), np.float32))
    assert _rel(mu.cpu().numpy(), g["mu"]) < 2e-5
    assert _rel(std.cpu().numpy(), g["std"]) < 2e-5
    assert _rel(z0.cpu().numpy(), g["mu"]) < 2e-5
    eps = np.random.default_rng(1).standard_normal((3, 2048)).astype(np.float32)
    z0, _, _ = engine.encode(0, g["enc_in"], eps)
    assert _rel(z0.cpu().numpy(), g["mu"] + eps * g["std"]) < 2e-5


def test_transforms_merge_smooth_match_oracle(engine, clip58):
    from oracle import pipeline_np as pl
    cams = clip58["camera_pose_list"]
    est = clip58["estimated_local_skeleton"]
    starts = pl.window_starts(len(est))
    pw = np.stack([est[s:s + 10] for s in starts])
    cw = np.stack([cams[s:s + 10] for s in starts])
    rel64, rel32 = engine.relative_global(torch.from_numpy(pw), torch.from_numpy(cw))
    ref = np.stack([pl.relative_global_pose(p, c) for p, c in zip(pw, cw)])
    np.testing.assert_allclose(rel64.cpu().numpy(), ref, rtol=0, atol=1e-13)
    assert (rel32.cpu().numpy() == ref.astype(np.float32)).mean() > 0.999
    # fp32 input (a stage result) is promoted to float64 exactly like numpy does
    rel64b, _ = engine.relative_global(torch.from_numpy(pw.astype(np.float32)), torch.from_numpy(cw))
    refb = np.stack([pl.relative_global_pose(p.astype(np.float32), c) for p, c in zip(pw, cw)])
    np.testing.assert_allclose(rel64b.cpu().numpy(), refb, rtol=0, atol=1e-13)
    glob = engine.to_global(rel64, torch.from_numpy(cw)).cpu().numpy()
    refg = np.stack([pl.to_global_pose(p, c) for p, c in zip(ref, cw)])
    np.testing.assert_allclose(glob, refg, rtol=0, atol=1e-13)
    merged = engine.merge_windows(torch.from_numpy(refg)).cpu().numpy()
    np.testing.assert_array_equal(merged, pl.merge_batches(refg))
    assert merged.shape == (8 * len(starts) + 2, 15, 3)
    sm = engine.gaussian_smooth(torch.from_numpy(merged)).cpu().numpy()
    from scipy.ndimage import gaussian_filter1d
    np.testing.assert_allclose(sm, gaussian_filter1d(merged, sigma=1, axis=0), rtol=0, atol=1e-14)
    one = engine.merge_windows(torch.from_numpy(refg[:1])).cpu().numpy()
    np.testing.assert_array_equal(one, refg[0])
    short = engine.gaussian_smooth(torch.from_numpy(merged[:3])).cpu().numpy()       # shorter than the kernel radius
    np.testing.assert_allclose(short, gaussian_filter1d(merged[:3], sigma=1, axis=0), rtol=0, atol=1e-14)
