"""The backward chain kernel's energy prologue (gemm_tap_tc.cu: Chain2Energy) against the stand-alone energy kernel:
both evaluate energy_device.cuh with the same reduction tree, so whole stage solves are bit-identical with the
prologue on (default) and off, for resident and zero-copy heat maps and ragged window counts.
Needs a B200: `pytest -m gpu`."""
import ctypes as C

import numpy as np
import pytest
import torch

from globalegomocap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

W_LOCAL = (0.01 / 10000, 0.001 / 100, 0.01, 0.0, 0.01)
W_GLOBAL = (0.01, 0.001, 0.01, 0.0, 0)
W_ALL = (0.01, 0.02, 0.05, 0.003, 0.04)


def _fuse(eng, on):
    eng.lib.gem_debug_fuse_energy.argtypes = [C.c_void_p, C.c_int]
    assert eng.lib.gem_debug_fuse_energy(eng._ctx, int(on)) == 0


@pytest.mark.parametrize("W", [1, 5, 12, 13, 24, 25, 100, 205])
def test_stage_solve_is_bit_identical_with_and_without_the_prologue(vae_weights, camera, W):
    from globalegomocap_b200.engine import Engine, energy_weights, lbfgs_params
    clip = syn.make_clip(8 * (W - 1) + 10, seed=80)
    est, heat = clip["estimated_local_skeleton"], clip["heatmap_list"]
    starts = syn.window_starts(len(est))
    assert len(starts) == W
    x0 = np.stack([est[s:s + 10] for s in starts]).astype(np.float32)
    mb = syn.mean_bone_length(est)
    eps = np.random.default_rng(W).standard_normal((W, 2048)).astype(np.float32)
    eng = Engine(max_windows=W)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    eng.set_vae(1, vae_weights[1])
    got = {}
    for which, wts, h in ((0, W_LOCAL, heat), (0, W_ALL, heat), (1, W_GLOBAL, None)):
        for on in (1, 0):
            _fuse(eng, on)
            n0 = eng.launch_count()
            r = eng.solve_stage(which, x0, h, None if h is None else np.asarray(starts, np.int64), np.zeros(W, np.int32), mb,
                                eps, energy_weights(*wts), lbfgs_params(max_iter=4), want_trace=True)
            torch.cuda.synchronize()
            got[on] = (r["pose"].clone(), r["trace"].clone(), r["n_iter"].clone(), r["func_evals"].clone(), r["status"].clone(),
                       eng.launch_count() - n0)
        for a, b in zip(got[1][:5], got[0][:5]):
            assert torch.equal(a, b) or (a.is_floating_point() and torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)))
        assert got[1][5] < got[0][5]                       # one launch per round fewer
    _fuse(eng, 1)
    eng.close()


def test_prologue_with_zero_copy_heat_maps(vae_weights, camera):
    """Texel cache + prefetch kernel in front of the fused chain (maps in pinned host memory)."""
    from globalegomocap_b200.engine import Engine, energy_weights, lbfgs_params
    W = 30
    clip = syn.make_clip(8 * (W - 1) + 10, seed=81)
    est = clip["estimated_local_skeleton"]
    starts = syn.window_starts(len(est))
    x0 = np.stack([est[s:s + 10] for s in starts]).astype(np.float32)
    mb = syn.mean_bone_length(est)
    eps = np.random.default_rng(3).standard_normal((W, 2048)).astype(np.float32)
    host_heat = torch.from_numpy(clip["heatmap_list"]).pin_memory()
    eng = Engine(max_windows=W)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    args = (0, x0, None, np.asarray(starts, np.int64), np.zeros(W, np.int32), mb, eps, energy_weights(*W_LOCAL),
            lbfgs_params(max_iter=3))
    out = {}
    for name, h, on in (("resident fused", clip["heatmap_list"], 1), ("zero-copy fused", host_heat, 1),
                        ("zero-copy unfused", host_heat, 0)):
        _fuse(eng, on)
        r = eng.solve_stage(args[0], args[1], h, *args[3:], want_trace=True)
        torch.cuda.synchronize()
        out[name] = (r["pose"].clone(), torch.nan_to_num(r["trace"], nan=-7.0))
    for name in ("zero-copy fused", "zero-copy unfused"):
        assert torch.equal(out[name][0], out["resident fused"][0]) and torch.equal(out[name][1], out["resident fused"][1]), name
    _fuse(eng, 1)
    eng.close()
