"""oracle/vae_np.py pinned against the reference ConvVAE (tests/golden/vae.npz)."""
import os

import numpy as np

from oracle.vae_np import VaeNp


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_decode_encode_vjp_match_reference(golden_dir, vae_weights):
    g = np.load(os.path.join(golden_dir, "vae.npz"))
    for dtype, tol in ((np.float32, 2e-5), (np.float64, 2e-5)):
        vae = VaeNp(vae_weights[0], dtype)
        pose, saved = vae.decode(g["z"], keep=True)
        assert pose.shape == (3, 10, 15, 3)
        assert _rel(pose, g["pose"]) < tol
        dz = vae.decode_vjp(saved, g["upstream"])
        assert _rel(dz, g["dz"]) < 10 * tol
        mu, std = vae.encode(g["enc_in"])
        assert _rel(mu, g["mu"]) < tol
        assert _rel(std, g["std"]) < tol


def test_vjp_is_adjoint_of_finite_difference(vae_weights):
    vae = VaeNp(vae_weights[1], np.float64)
    rng = np.random.default_rng(1)
    z = rng.standard_normal((1, 2048))
    dzv = rng.standard_normal((1, 2048))
    up = rng.standard_normal((1, 10, 15, 3))
    pose, saved = vae.decode(z, keep=True)
    h = 1e-6
    fd = ((vae.decode(z + h * dzv) - vae.decode(z - h * dzv)) / (2 * h) * up).sum()
    an = (vae.decode_vjp(saved, up) * dzv).sum()
    assert abs(fd - an) < 1e-6 * max(1.0, abs(fd))
