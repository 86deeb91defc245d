"""Planar ([frames, J, H, W]) and tiled ([frames, J, H/4, W/8, 4, 8]) heat maps (gem_ctx_set_heat_layout) against the
pickle's HWC layout: the same texel values reach the same arithmetic, so energies, gradients and whole solves are
bit-identical; over PCIe the planar window needs fewer requests than the HWC one, the tiled one fewer again.
Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest
import torch

from globalegomocap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

W_LOCAL = (0.01 / 10000, 0.001 / 100, 0.01, 0.0, 0.01)


def test_energy_and_gradient_are_bit_identical_in_both_layouts(golden_dir, clip58, camera):
    import os
    from globalegomocap_b200.engine import Engine, energy_weights, tile_heat
    g = np.load(os.path.join(golden_dir, "energy.npz"))
    names = [str(n) for n in g["names"] if str(n).startswith("all__")]
    xs = np.stack([g[f"{n}__x"] for n in names])
    starts = [int(g[f"{n}__start"]) for n in names]
    x0 = np.stack([clip58["estimated_local_skeleton"][s:s + 10] for s in starts]).astype(np.float32)
    heat = np.concatenate([syn.dense_heat_window(1) if (n.endswith("edges") or n.endswith("dense_near"))
                           else clip58["heatmap_list"][s:s + 10] for n, s in zip(names, starts)])
    fb = np.arange(len(names), dtype=np.int64) * 10
    eng = Engine(max_windows=16)
    eng.set_camera(*camera)
    w = energy_weights(*g["all__weights"])
    args = (xs, x0, None, fb, np.zeros(len(names), np.int32), g["mean_bone_length"], w)
    a = eng.energy_grad(args[0], args[1], heat, *args[3:])
    eng.set_heat_layout(True)
    planar = np.ascontiguousarray(heat.transpose(0, 3, 1, 2))
    b = eng.energy_grad(args[0], args[1], planar, *args[3:])
    eng.set_heat_layout("tiled")            # [frames, J, H/4, W/8, 4, 8]: the same texels behind another address
    c = eng.energy_grad(args[0], args[1], tile_heat(torch.from_numpy(planar)).contiguous(), *args[3:])
    for u, v, t in zip(a, b, c):
        assert torch.equal(u, v) and torch.equal(u, t)
    with pytest.raises(Exception):          # planar-shaped maps while the engine expects tiles
        eng.energy_grad(args[0], args[1], planar, *args[3:])
    eng.close()


def test_solves_are_bit_identical_and_pcie_requests_drop(vae_weights, camera):
    from globalegomocap_b200.engine import Engine, energy_weights, lbfgs_params
    W = 120
    clip = syn.make_clip(8 * (W - 1) + 10, seed=90)
    est = clip["estimated_local_skeleton"]
    starts = syn.window_starts(len(est))
    x0 = np.stack([est[s:s + 10] for s in starts]).astype(np.float32)
    mb = syn.mean_bone_length(est)
    eps = np.random.default_rng(9).standard_normal((W, 2048)).astype(np.float32)
    hwc = torch.from_numpy(np.ascontiguousarray(clip["heatmap_list"]))
    planar = hwc.permute(0, 3, 1, 2).contiguous()

    def pinned(t):
        p = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        p.copy_(t)
        assert p.is_pinned() and p.is_contiguous()
        return p

    eng = Engine(max_windows=W)
    eng.set_camera(*camera)
    eng.set_vae(0, vae_weights[0])
    out, fetched, looked = {}, {}, {}
    from globalegomocap_b200.engine import tile_heat
    tiled = tile_heat(planar).contiguous()
    for name, heat, is_planar in (("hwc resident", hwc.cuda(), False), ("planar resident", planar.cuda(), True),
                                  ("hwc zero-copy", pinned(hwc), False), ("planar zero-copy", pinned(planar), True),
                                  ("planar resident, cache forced", planar.cuda(), True),
                                  ("tiled resident", tiled.cuda(), 2), ("tiled zero-copy", pinned(tiled), 2),
                                  ("tiled resident, cache forced", tiled.cuda(), 2)):
        eng.set_heat_layout(is_planar)
        eng.set_texel_cache(1 if "forced" in name else -1)        # forced: the energy kernel's own per-thread row fetch
        eng.texel_cache_stats(True)
        r = eng.solve_stage(0, x0, heat, np.asarray(starts, np.int64), np.zeros(W, np.int32), mb, eps, energy_weights(*W_LOCAL),
                            lbfgs_params(max_iter=6), want_trace=True)
        torch.cuda.synchronize()
        looked[name], fetched[name] = eng.texel_cache_stats(False)
        out[name] = (r["pose"].clone(), torch.nan_to_num(r["trace"], nan=-7.0), r["func_evals"].clone())
    for name in out:
        for u, v in zip(out[name], out["hwc resident"]):
            assert torch.equal(u, v), name
    print("cache lookups:", looked, "PCIe requests:", fetched)
    assert looked["hwc resident"] == looked["planar resident"] == 0           # small batches of resident maps: no cache
    assert looked["hwc zero-copy"] == looked["planar zero-copy"] > 0
    assert looked["planar resident, cache forced"] == looked["planar zero-copy"]
    # HWC: one request per texel.  Planar: whole window rows (16 texels = two 32-byte sectors, ONE request each; the
    # counter is in sectors), and a miss also brings in the rows next to the footprint's
    assert 0 < fetched["planar zero-copy"] / 2 < 0.8 * fetched["hwc zero-copy"]
    # tiled: one request = one 4 x 8 tile (four sectors): fewer requests again
    assert looked["tiled resident"] == 0 and looked["tiled zero-copy"] == looked["tiled resident, cache forced"] == looked["planar zero-copy"]
    assert 0 < fetched["tiled zero-copy"] / 4 < 0.8 * fetched["planar zero-copy"] / 2
    eng.close()
