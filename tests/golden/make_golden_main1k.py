#!/usr/bin/env python
"""BASELINE configs[0] as a golden: the UNMODIFIED reference's `optimizer.main` (optimizer.py:311-507) on one synthetic
1000-frame sequence (124 windows, both stages, the reference's own max_iter = 25), run in this CPU container through
the harness of make_golden.py (open3d / natsort stubbed, random-init checkpoints in a scratch CWD, injected
reparameterisation noise).

The 245 MB of inputs are not stored: `synthetic.make_clip(1000, seed=SEED)` regenerates them bit for bit, and the noise
comes from `numpy.random.default_rng(EPS_SEED)`.  Stored: the four returned sequences and the 18 error metrics
(tests/golden/main_1k_mi25.npz, ~1 MB).

    python tests/golden/make_golden_main1k.py [--threads 8]
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

syn = mg.syn
SEED, EPS_SEED, N_FRAMES = 23, 5151, 1000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=8)
    args = ap.parse_args()
    import torch
    clip = syn.make_clip(N_FRAMES, seed=SEED)
    # the checkpoints of every golden (and of tests/conftest.py's `vae_weights`): seeds 11 / 12 with the mean-pose bias
    # of make_golden.py's 58-frame clip
    h = mg.Harness(tempfile.mkdtemp(prefix="gem_golden_1k_"), syn.make_clip(58, seed=7))
    inj = mg.EpsInjector(h.ref)
    data_dir = os.path.join(h.scratch, "data", "synth", "clip1k")
    syn.write_clip_pickle(clip, data_dir)
    W = len(syn.window_starts(N_FRAMES))
    eps = np.random.default_rng(EPS_SEED).standard_normal((W, 2, 2048)).astype(np.float32)

    def run(threads):
        torch.set_num_threads(threads)
        for w in range(W):                  # call order: local(w0), global(w0), local(w1), ...
            inj.push(eps[w, 0])
            inj.push(eps[w, 1])
        t0 = time.perf_counter()
        res = h.ref.main(data_dir, camera_model_path=h.camera_json, vae_weight=0.0, gmm_weight=0.0,
                         smoothness_weight=0.001, bone_length_weight=0.01, weight_3d=0.01, reproj_weight=0.01,
                         visualization=False, save=False, merge=True, final_smooth=True)
        return res, time.perf_counter() - t0

    try:
        (errors, est_seq, mid_local, opt_seq, gt_seq), seconds = run(args.threads)
        # the same call at ONE thread: the reference's own run-to-run noise on this workload (the yardstick the test
        # prints next to the CUDA path's deviation)
        (errors1, _, mid_local1, opt_seq1, _), seconds1 = run(1)
    finally:
        inj.restore()
    out = {"seed": np.int64(SEED), "eps_seed": np.int64(EPS_SEED), "n_frames": np.int64(N_FRAMES), "windows": np.int64(W),
           "max_iter": np.int64(25), "threads": np.int64(args.threads), "reference_seconds": np.float64(seconds),
           "reference_seconds_one_thread": np.float64(seconds1),
           "final_estimated_seq": np.asarray(est_seq), "mid_local_pose_seq": np.asarray(mid_local, dtype=np.float32),
           "final_optimized_seq": np.asarray(opt_seq), "final_gt_seq": np.asarray(gt_seq),
           "mid_local_pose_seq_one_thread": np.asarray(mid_local1, dtype=np.float32),
           "final_optimized_seq_one_thread": np.asarray(opt_seq1)}
    for k, v in errors.items():
        out["err__" + k] = np.asarray(v)
        out["err1__" + k] = np.asarray(errors1[k])
    np.savez_compressed(os.path.join(mg.OUT, "main_1k_mi25.npz"), **out)
    print("reference main on %d frames (%d windows): %.1f s at %d threads = %.2f frames/s; %.1f s at one thread" %
          (N_FRAMES, W, seconds, args.threads, 8 * W / seconds, seconds1))
    print({k: float(np.mean(v)) for k, v in errors.items()})


if __name__ == "__main__":
    main()
