#!/usr/bin/env python
"""Runs the UNMODIFIED reference (baseline/_ref, see baseline/install_reference.py) over the windows of
tests/golden/dist.npz on the device torch offers it — on the GPU box that is the reference's own CUDA path
(optimizer.py:39: cuBLAS / cuDNN) — and writes per-evaluation energies and final poses in dist.npz's layout.
tests/test_gpu_dist.py starts it as a subprocess: "reference on CUDA vs reference on one CPU thread" is the
reference's own cross-backend noise, the yardstick for the CUDA path's deviations.

    python tests/reference_runs.py --out /tmp/ref_cuda.npz [--tf32 0|1] [--max-iters 3 25] [--windows 64]
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--tf32", type=int, default=-1, help="-1: torch's defaults (cuDNN TF32 allowed), 0: plain fp32 everywhere, 1: TF32 everywhere")
    ap.add_argument("--max-iters", type=int, nargs="*", default=[3, 25])
    ap.add_argument("--windows", type=int, default=64)
    args = ap.parse_args()
    import torch
    from baseline import reference_harness as rh
    from globalegomocap_b200 import synthetic as syn
    root = rh.reference_root()
    if root is None:
        print("reference not installed (baseline/_ref)")
        return 3
    if args.tf32 >= 0:
        torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
        torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    g = np.load(os.path.join(REPO, "tests", "golden", "dist.npz"))
    clip = syn.make_clip(int(g["n_frames"]), seed=int(g["clip_seed"]))
    bias = syn.mean_pose_bias(clip)
    weights = (syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias), syn.make_vae_state_dict(12, perturb_bn=True, pose_bias=bias))
    starts = [int(s) for s in g["starts"]][:args.windows]
    W = len(starts)
    out = {"device": np.asarray("cuda" if torch.cuda.is_available() else "cpu"), "tf32": np.int64(args.tf32)}
    for mi in args.max_iters:
        run = rh.ReferenceRunner(root, clip, weights, syn.DEFAULT_CAMERA_JSON, max_iter=mi)
        E = np.full((2, W, 34), np.nan)
        pose = np.zeros((2, W, 10, 15, 3), np.float32)
        n_eval = np.zeros((2, W), np.int32)
        try:
            for wi, s in enumerate(starts):
                for stage in (0, 1):
                    x0 = clip["estimated_local_skeleton"][s:s + 10] if stage == 0 else g[f"mi{mi}_rel_in"][wi]
                    p, en = run.solve_stage(stage, x0, clip["heatmap_list"][s:s + 10], g["eps"][wi, stage])
                    E[stage, wi, :len(en)] = en
                    pose[stage, wi] = p
                    n_eval[stage, wi] = len(en)
        finally:
            run.close()
        out[f"mi{mi}_E"], out[f"mi{mi}_pose"], out[f"mi{mi}_n_eval"] = E, pose, n_eval
    np.savez_compressed(args.out, **out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
