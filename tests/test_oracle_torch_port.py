"""oracle/torch_port.py (the timed CPU baseline) reproduces the reference's traces."""
import os
import time

import numpy as np
import torch

from oracle import energy_np as en
from oracle.torch_port import TorchPortOptimizer

W_LOCAL = (0.01 / 10000, 0.001 / 100, 0.01, 0.0, 0.01)
W_GLOBAL = (0.01, 0.001, 0.01, 0.0, 0)


def test_torch_port_tracks_reference_25_iterations(golden_dir, clip58, vae_weights, camera):
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, "traces.npz"))
    mb = en.mean_bone_length(clip58["estimated_local_skeleton"])
    wi, s = 1, int(g["starts"][1])
    heat = clip58["heatmap_list"][s:s + 10]
    opt = TorchPortOptimizer(vae_weights[0], camera, mb, W_LOCAL, max_iter=25)
    trace = []
    t0 = time.perf_counter()
    pose, info = opt.solve(clip58["estimated_local_skeleton"][s:s + 10], heat, g["eps"][wi, 0], trace)
    dt = time.perf_counter() - t0
    E_ref = g[f"mi25_w{wi}_local_E"]
    assert info["func_evals"] == len(E_ref) == len(trace)
    E = np.array([t[0] for t in trace])
    # same ATen ops, same thread count as the golden run: bit-for-bit in practice
    assert np.abs(E - E_ref).max() <= 1e-6 * np.abs(E_ref).max()
    assert np.abs(pose - g[f"mi25_w{wi}_local_pose"]).max() * 1000 < 0.01
    assert dt < 60
    for stage, W, sd, x0, e in (("global", W_GLOBAL, vae_weights[1], g[f"mi3_w{wi}_local_relglobal"], 1),):
        opt = TorchPortOptimizer(sd, camera, mb, W, max_iter=3)
        trace = []
        pose, info = opt.solve(x0, heat, g["eps"][wi, e], trace)
        E_ref = g[f"mi3_w{wi}_{stage}_E"]
        assert len(trace) == len(E_ref)
        assert np.abs(np.array([t[0] for t in trace]) - E_ref).max() <= 1e-6 * np.abs(E_ref).max()
