"""Host-side weight preparation (BatchNorm folding, tap arrangement, Linear+ConvT fusion,
fc permutation) checked on CPU: the prepared operands, pushed through a plain-torch model of
the tap GEMM, must reproduce the oracle's layer-by-layer VAE."""
import os

import numpy as np
import torch

from globalegomocap_b200.vae_prep import PreparedVae
from oracle.vae_np import VaeNp


def tap_gemm_ref(A, layer, T, epi=None, aux=None):
    """out[m] = bias + sum_tap A[m + tap - taps//2] @ W[tap], taps leaving the window read zeros."""
    M = A.shape[0]
    w = layer["w"].double()[:, :, :layer["n"]]
    out = torch.zeros(M, layer["n"], dtype=torch.float64)
    t_idx = torch.arange(M) % T
    for tap in range(layer["taps"]):
        shift = tap - layer["taps"] // 2
        ok = ((t_idx + shift) >= 0) & ((t_idx + shift) < T) if layer["taps"] > 1 else torch.ones(M, dtype=torch.bool)
        src = torch.zeros_like(A, dtype=torch.float64)
        rows = torch.nonzero(ok).squeeze(1)
        src[rows] = A.double()[rows + shift]
        out += src @ w[tap]
    if layer["bias"] is not None:
        out += layer["bias"].double()[None, :]
    if epi == "lrelu":
        out = torch.where(out > 0, out, out * 0.01)
    elif epi == "mask":
        out = torch.where(aux > 0, out, out * 0.01)
    return out


def decode_ref(p, z):
    T, W = p.seq_len, z.shape[0]
    acts = []
    h = tap_gemm_ref(z, p.dec[0], T, "lrelu").reshape(W * T, -1)
    acts.append(h)
    for i in range(1, 5):
        h = tap_gemm_ref(h, p.dec[i], T, "lrelu")
        acts.append(h)
    pose = tap_gemm_ref(h, p.dec[5], T)
    return pose.reshape(W, T, 15, 3), acts


def decode_vjp_ref(p, acts, dpose):
    T, W = p.seq_len, dpose.shape[0]
    g = dpose.reshape(W * T, -1)
    for i in range(5):
        g = tap_gemm_ref(g, p.dec_bwd[i], T, "mask", acts[4 - i])
    return tap_gemm_ref(g.reshape(W, -1), p.dec_bwd[5], T)


def encode_ref(p, pose):
    T, W = p.seq_len, pose.shape[0]
    h = pose.reshape(W * T, -1)
    for i in range(5):
        h = tap_gemm_ref(h, p.enc[i], T, "lrelu")
    fc = tap_gemm_ref(h.reshape(W, -1), p.enc[5], T)
    n = p.latent_dim
    return fc[:, :n], torch.exp(0.5 * fc[:, n:])


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / np.abs(b).max()


def test_prepared_operands_reproduce_the_vae(golden_dir, vae_weights):
    g = np.load(os.path.join(golden_dir, "vae.npz"))
    p = PreparedVae(vae_weights[0], "cpu")
    assert [(l["taps"], l["k"], l["n"]) for l in p.dec] == [(1, 2048, 2560), (3, 256, 128), (3, 128, 64), (3, 64, 64),
                                                           (3, 64, 64), (3, 64, 45)]
    assert [(l["taps"], l["k"], l["n"]) for l in p.dec_bwd] == [(3, 45, 64), (3, 64, 64), (3, 64, 64), (3, 64, 128),
                                                               (3, 128, 256), (1, 2560, 2048)]
    assert p.dec[5]["w"].shape == (3, 64, 48)          # rows padded to a multiple of 4 floats
    ref = VaeNp(vae_weights[0], np.float64)
    pose, acts = decode_ref(p, torch.from_numpy(g["z"]).double())
    pose_o, saved = ref.decode(g["z"].astype(np.float64), keep=True)
    assert _rel(pose.numpy(), pose_o) < 2e-6
    assert _rel(pose.numpy(), g["pose"]) < 2e-5        # and the reference's own output
    dz = decode_vjp_ref(p, acts, torch.from_numpy(g["upstream"]).double())
    assert _rel(dz.numpy(), ref.decode_vjp(saved, g["upstream"].astype(np.float64))) < 2e-6
    assert _rel(dz.numpy(), g["dz"]) < 2e-4
    mu, std = encode_ref(p, torch.from_numpy(g["enc_in"]).double())
    assert _rel(mu.numpy(), g["mu"]) < 2e-5 and _rel(std.numpy(), g["std"]) < 2e-5
