"""oracle/energy_np.py pinned against the unmodified reference's energy terms
and autograd gradients (tests/golden/energy.npz, fisheye.npz)."""
import os

import numpy as np
import pytest

from globalegomocap_b200 import synthetic as syn
from oracle import energy_np as en

TERMS = ("e3d", "smooth", "bone", "vae", "reproj")


def _heat_for(name, clip58, start):
    if name.endswith("edges") or name.endswith("dense_near"):
        return syn.dense_heat_window(1)
    return clip58["heatmap_list"][start:start + 10]


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("tag", ["default", "new"])
def test_fisheye_projection_and_jacobian(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "fisheye.npz"))
    cam = syn.load_camera(syn.DEFAULT_CAMERA_JSON if tag == "default"
                          else syn.DEFAULT_CAMERA_JSON.replace("calibration.json", "calibration_new.json"))
    pts = g[f"{tag}_points"]
    uv, aux = en.fisheye_project(pts, *cam)
    np.testing.assert_allclose(uv, g[f"{tag}_uv"], rtol=2e-6, atol=2e-4)
    J = en.fisheye_jacobian(pts, aux)
    assert _rel(J[:, 0, :], g[f"{tag}_du"]) < 2e-5
    assert _rel(J[:, 1, :], g[f"{tag}_dv"]) < 2e-5
    if tag == "default":        # SURVEY.md A.5 sanity values
        np.testing.assert_allclose(uv[0], [695.3890, 601.4163], atol=2e-3)
        np.testing.assert_allclose(uv[1], [502.7032, 582.3908], atol=2e-3)


def test_norm_is_zero_raises(camera):
    with pytest.raises(Exception, match="norm is zero!"):
        en.fisheye_project(np.array([[0.0, 0.0, 1.0]], dtype=np.float32), *camera)


def test_terms_and_gradients_match_reference(golden_dir, clip58, camera):
    g = np.load(os.path.join(golden_dir, "energy.npz"))
    mb = g["mean_bone_length"]
    np.testing.assert_allclose(en.mean_bone_length(clip58["estimated_local_skeleton"]), mb, rtol=1e-6)
    for name in g["names"]:
        name = str(name)
        start = int(g[f"{name}__start"])
        x = g[f"{name}__x"]
        x0 = clip58["estimated_local_skeleton"][start:start + 10].astype(np.float32)
        heat = _heat_for(name, clip58, start)
        wname = name.split("__")[0]
        weights = tuple(g[f"{wname}__weights"])
        for dtype, tol_e, tol_g in ((np.float32, 3e-6, 3e-5), (np.float64, 3e-6, 3e-5)):
            if dtype is np.float64 and name.endswith("edges"):
                continue        # integer-valued pixel coordinates: floor() flips between fp32 and fp64
            xd = x.astype(dtype)
            res = {"e3d": en.e3d(xd, x0.astype(dtype)), "smooth": en.esmooth(xd), "bone": en.ebone(xd, mb),
                   "vae": en.evae(xd), "reproj": en.ereproj(xd, heat, *camera)}
            for t in TERMS:
                e_ref, g_ref = float(g[f"{name}__E_{t}"]), g[f"{name}__G_{t}"]
                assert abs(res[t][0] - e_ref) <= tol_e * max(abs(e_ref), 1.0), (name, t, dtype)
                assert _rel(res[t][1], g_ref) < tol_g, (name, t, dtype, _rel(res[t][1], g_ref))
            E, G, _ = en.total_energy(xd, x0, heat, mb, weights, *camera)
            e_ref = float(g[f"{name}__E_total"])
            assert abs(E - e_ref) <= 1e-5 * max(abs(e_ref), 1e-2), (name, dtype)
            assert _rel(G, g[f"{name}__G_total"]) < 3e-5, (name, dtype)


def test_border_samples_keep_partial_weights(camera):
    """Zero padding keeps the in-range corners' weights just outside the map
    (SURVEY.md A.5): at ix = -0.3 the sample is 0.7*H[y,0] and dS/dix = +H."""
    maps = np.ones((1, 64, 64), dtype=np.float32)
    S, dx, dy = en.bilinear_zeros(maps, np.array([-0.3], np.float32), np.array([20.0], np.float32))
    np.testing.assert_allclose(S, [0.7], rtol=1e-6)
    np.testing.assert_allclose(dx, [1.0], rtol=1e-6)
    S, dx, dy = en.bilinear_zeros(maps, np.array([63.0], np.float32), np.array([20.0], np.float32))
    np.testing.assert_allclose(S, [1.0], rtol=1e-6)
    np.testing.assert_allclose(dx, [-1.0], rtol=1e-6)          # x1 = 64 is outside: its texel reads 0
    S, dx, dy = en.bilinear_zeros(maps, np.array([-1.01], np.float32), np.array([20.0], np.float32))
    assert S[0] == 0 and dx[0] == 0


def test_gradient_finite_difference(clip58, camera):
    rng = np.random.default_rng(0)
    x = clip58["estimated_local_skeleton"][:10] + 0.003 * rng.standard_normal((10, 15, 3))
    x0 = clip58["estimated_local_skeleton"][:10]
    heat = syn.dense_heat_window(2)
    mb = en.mean_bone_length(clip58["estimated_local_skeleton"])
    w = (0.01, 0.02, 0.05, 0.003, 0.04)
    E, G, _ = en.total_energy(x, x0, heat, mb, w, *camera)
    d = rng.standard_normal(x.shape)
    h = 1e-7
    Ep, _, _ = en.total_energy(x + h * d, x0, heat, mb, w, *camera)
    Em, _, _ = en.total_energy(x - h * d, x0, heat, mb, w, *camera)
    fd = (Ep - Em) / (2 * h)
    assert abs(fd - (G * d).sum()) < 1e-5 * max(1.0, abs(fd))
