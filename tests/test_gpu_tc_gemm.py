"""tcgen05/TMEM/TMA 3xTF32 GEMM path against the fp32 CUDA-core path and the reference goldens.
Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(scope="module")
def engine(vae_weights, camera):
    from globalegomocap_b200.engine import Engine
    eng = Engine(max_windows=700)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    eng.set_vae(1, vae_weights[1])
    yield eng
    eng.close()


@pytest.mark.parametrize("M,N,K", [(64, 128, 32), (128, 128, 64), (64, 128, 96), (64, 128, 128), (64, 128, 2560),
                                   (64, 2048, 2560), (300, 2560, 2048), (641, 256, 5120)])
def test_gemm_kernels_against_float64(engine, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(K, N, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref = a.double() @ b.double() + bias.double()
    ref_act = torch.where(ref > 0, ref, ref * 0.01)
    for tc in (False, True):
        c = engine.gemm(a, b, bias, leaky_relu=False, tensor_cores=tc).cpu()
        c_act = engine.gemm(a, b, bias, leaky_relu=True, tensor_cores=tc).cpu()
        r, r_act = _rel(c, ref), _rel(c_act, ref_act)
        print((M, N, K), "tensor cores" if tc else "cuda cores", "relative max error vs float64:", r, r_act)
        assert r < 1e-5 and r_act < 1e-5, (M, N, K, tc, r, r_act)


@pytest.mark.parametrize("W", [64, 128, 300, 641])
def test_tc_gemm_matches_fp32_path(engine, W):
    """Full tiles, a ragged last M tile, several waves: decoder fwd, bwd-data and encoder."""
    g = torch.Generator(device="cpu").manual_seed(W)
    z = torch.randn(W, 2048, generator=g)
    up = torch.randn(W, 10, 15, 3, generator=g)
    pose_in = torch.randn(W, 10, 45, generator=g) * 0.3
    eps = torch.randn(W, 2048, generator=g)
    out = {}
    engine.set_gemm_mode(0)
    engine.decode(0, z)          # activations (LeakyReLU masks) of this decode are used by BOTH vjp calls:
    for mode in (0, 1):          # a mask that flips between two 1e-6-different forwards would dominate dz
        engine.set_gemm_mode(mode)
        dz = engine.decode_vjp(0, up)
        pose = engine.decode(0, z)
        z0, mu, std = engine.encode(1, pose_in, eps)
        torch.cuda.synchronize()
        out[mode] = (pose.clone(), dz.clone(), z0.clone(), mu.clone(), std.clone())
        engine.set_gemm_mode(0)
        engine.decode(0, z)
    engine.set_gemm_mode(0)
    names = ("pose", "dz", "z0", "mu", "std")
    for name, a, b in zip(names, out[1], out[0]):
        r = _rel(a, b)
        print(W, name, "tcgen05 vs fp32 CUDA-core relative max error:", r)
        # mu: K = 5120 and cancellation in the sum; dz: six tensor-core layers in a row, each adding the
        # tensor core's truncating fp32 accumulation (~1e-6 per layer, gemm_tc.cu)
        assert r < {"mu": 3e-5, "dz": 2e-5}.get(name, 1e-5), (W, name, r)


def test_tc_gemm_matches_reference_goldens(engine, golden_dir):
    """M < 64 falls back to the CUDA-core kernel by design; replicate the golden rows to fill a tile."""
    g = np.load(os.path.join(golden_dir, "vae.npz"))
    reps = 43
    z = np.tile(g["z"], (reps, 1))
    up = np.tile(g["upstream"], (reps, 1, 1, 1))
    engine.set_gemm_mode(1)
    pose = engine.decode(0, z).cpu().numpy()
    dz = engine.decode_vjp(0, up).cpu().numpy()
    z0, mu, std = engine.encode(0, np.tile(g["enc_in"], (reps, 1, 1)), np.zeros((3 * reps, 2048), np.float32))
    engine.set_gemm_mode(0)
    for r in range(0, reps, 14):
        sl = slice(3 * r, 3 * r + 3)
        assert np.abs(pose[sl] - g["pose"]).max() / np.abs(g["pose"]).max() < 2e-5
        assert np.abs(dz[sl] - g["dz"]).max() / np.abs(g["dz"]).max() < 2e-4
        assert np.abs(mu.cpu().numpy()[sl] - g["mu"]).max() / np.abs(g["mu"]).max() < 2e-5


@pytest.mark.parametrize("W", [1, 2, 11, 12, 13, 37, 149])
def test_tc_tap_chain_ragged_window_counts(engine, W):
    """The tcgen05 tap kernel packs 3 windows per 32-lane TMEM quarter (12 per CTA): window counts that
    leave quarters / tiles partly empty, and a single window, must match the CUDA-core layers."""
    g = torch.Generator(device="cpu").manual_seed(1000 + W)
    z = torch.randn(W, 2048, generator=g)
    up = torch.randn(W, 10, 15, 3, generator=g)
    engine.set_gemm_mode(1)
    pose_tc = engine.decode(0, z).clone()          # leaves the (split) activations both vjp calls mask with
    dz_tc = engine.decode_vjp(0, up).clone()
    engine.set_gemm_mode(0)
    dz_simt = engine.decode_vjp(0, up).clone()
    pose_simt = engine.decode(0, z).clone()
    torch.cuda.synchronize()
    assert _rel(pose_tc, pose_simt) < 1e-5, (W, _rel(pose_tc, pose_simt))
    assert _rel(dz_tc, dz_simt) < 2e-5, (W, _rel(dz_tc, dz_simt))   # six tensor-core layers in a row


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (300, 256, 128), (1870, 2560, 2048), (641, 2048, 2560), (130, 4096, 5120)])
def test_fp16_scheme_gemm_against_float64(engine, M, N, K):
    """tensor_cores=2: x ~ fp16 hi + 2^-11 fp16 lo, three kind::f16 MMAs, cross terms in their own accumulator.
    Inputs with a spread of magnitudes (columns scaled by 1e-4 and 300) must come out at the same 1e-5 level
    as the 3xTF32 scheme."""
    g = torch.Generator(device="cpu").manual_seed(7 * M + N + K)
    a = torch.randn(M, K, generator=g)
    a[:, ::7] *= 1e-4
    a[:, 3::11] *= 300.0
    b = torch.randn(K, N, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref = a.double() @ b.double() + bias.double()
    errs = {}
    for mode in (1, 2):
        c = engine.gemm(a, b, bias, leaky_relu=False, tensor_cores=mode).cpu()
        errs[mode] = _rel(c, ref)
    print((M, N, K), "relative max error vs float64: 3xTF32", errs[1], " fp16 scheme", errs[2])
    assert errs[2] < 1e-5, errs


@pytest.mark.parametrize("W", [1, 13, 128, 300, 641])
def test_fp16_chain_matches_fp32_path(engine, W):
    """gemm_mode 3: every tensor-core layer in the fp16 scheme (fp16 hi + 2^-11 fp16 lo activations, three fp16 weight
    slabs, gradient rows rescaled by a power of two in the energy kernel / the entry split) against the CUDA-core
    layers, including gradients a million times smaller than the activations."""
    g = torch.Generator(device="cpu").manual_seed(3000 + W)
    z = torch.randn(W, 2048, generator=g)
    for scale in (1.0, 1e-6):
        up = torch.randn(W, 10, 15, 3, generator=g) * scale
        engine.set_gemm_mode(3)
        pose_tc = engine.decode(0, z).clone()
        dz_tc = engine.decode_vjp(0, up).clone()
        engine.set_gemm_mode(0)
        dz_simt = engine.decode_vjp(0, up).clone()      # same masks: the signs the mode-3 decode left
        pose_simt = engine.decode(0, z).clone()
        torch.cuda.synchronize()
        rp, rz = _rel(pose_tc, pose_simt), _rel(dz_tc, dz_simt)
        print(W, scale, "fp16 chain vs fp32 CUDA cores: pose", rp, "dz", rz)
        assert rp < 1e-5 and rz < 2e-5, (W, scale, rp, rz)
    engine.set_gemm_mode(2)


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (300, 256, 128), (1870, 2560, 2048), (641, 2048, 2560), (130, 4096, 5120),
                                   (257, 2560, 2048)])
def test_fp16_scheme_cta_pair_matches_single_cta(engine, M, N, K):
    """The CTA-pair kernel (tcgen05.mma.cta_group::2, M = 256, each CTA holding half of the B tile) runs the same
    MMAs in the same order per output element as the one-CTA kernel: bit-identical results, for odd M-tile counts
    (a pair whose second CTA is empty), ragged last rows and N tiles that stick out of the matrix."""
    import ctypes as C
    g = torch.Generator(device="cpu").manual_seed(11 * M + N + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(K, N, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref = a.double() @ b.double() + bias.double()
    out = {}
    lib = engine.lib
    lib.gem_debug_gemm_pair.argtypes = [C.c_int]
    try:
        for pair in (0, 1):
            lib.gem_debug_gemm_pair(pair)
            out[pair] = engine.gemm(a, b, bias, leaky_relu=True, tensor_cores=2).cpu()
    finally:
        lib.gem_debug_gemm_pair(-1)
    ref_act = torch.where(ref > 0, ref, ref * 0.01)
    assert _rel(out[1], ref_act) < 1e-5
    assert torch.equal(out[0], out[1]), float((out[0] - out[1]).abs().max())


@pytest.mark.parametrize("W", [1, 2, 11, 12, 13, 37, 149, 641])
def test_tap_chain_matches_layer_by_layer(engine, W):
    """gemm_mode 3 runs the k=3 layers of each direction as ONE launch (the activation tile stays in shared memory,
    weights stream through a ring): four layers per one-CTA chain (mode 1) or all five on CTA pairs with
    tcgen05.mma.cta_group::2 (mode 2, the default).  Same MMAs per output element, same epilogue arithmetic: pose
    and dz must be bit-identical to one launch per layer, for window counts that leave TMEM quarters, tiles or the
    second CTA of a pair partly or wholly empty."""
    import ctypes as C
    g = torch.Generator(device="cpu").manual_seed(5000 + W)
    z = torch.randn(W, 2048, generator=g)
    up = torch.randn(W, 10, 15, 3, generator=g) * 1e-3
    lib = engine.lib
    lib.gem_debug_tap_chain.argtypes = [C.c_void_p, C.c_int]
    out = {}
    engine.set_gemm_mode(3)
    try:
        for chain in (0, 1, 2):
            lib.gem_debug_tap_chain(engine._ctx, chain)
            pose = engine.decode(0, z).clone()
            dz = engine.decode_vjp(0, up).clone()
            torch.cuda.synchronize()
            out[chain] = (pose, dz)
    finally:
        lib.gem_debug_tap_chain(engine._ctx, 2)
        engine.set_gemm_mode(2)
    for chain in (1, 2):        # 1: one-CTA chains of four layers, 2: CTA-pair chains of five (the default)
        assert torch.isfinite(out[chain][0]).all() and torch.isfinite(out[chain][1]).all()
        assert torch.equal(out[0][0], out[chain][0]), (chain, float((out[0][0] - out[chain][0]).abs().max()))
        assert torch.equal(out[0][1], out[chain][1]), (chain, float((out[0][1] - out[chain][1]).abs().max()))


@pytest.mark.parametrize("W", [1, 13, 300, 641])
def test_fp16_encoder_matches_fp32_path(engine, W):
    """The encoder's 128 -> 256 -> 512 k=3 layers on the tcgen05 tap kernel (fp16 scheme; default, GEM_ENC_TC=0 opts out),
    their split output fed straight to the fc GEMM; mu / std / z0 must match the CUDA-core layers."""
    import ctypes as C
    g = torch.Generator(device="cpu").manual_seed(7000 + W)
    pose_in = torch.randn(W, 10, 45, generator=g) * 0.3
    eps = torch.randn(W, 2048, generator=g)
    out = {}
    engine.lib.gem_debug_enc_tc.argtypes = [C.c_void_p, C.c_int]
    engine.lib.gem_debug_enc_tc(engine._ctx, 1)         # the default: tensor-core encoder layers in mode 3
    try:
        for mode in (0, 3):
            engine.set_gemm_mode(mode)
            z0, mu, std = engine.encode(1, pose_in, eps)
            torch.cuda.synchronize()
            out[mode] = (z0.clone(), mu.clone(), std.clone())
    finally:
        engine.set_gemm_mode(2)
    for name, a, b in zip(("z0", "mu", "std"), out[3], out[0]):
        r = _rel(a, b)
        print(W, name, "fp16 encoder vs fp32 CUDA cores:", r)
        assert r < 3e-5, (W, name, r)
