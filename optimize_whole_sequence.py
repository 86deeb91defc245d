#!/usr/bin/env python
"""Drop-in replacement for the reference's `optimize_whole_sequence.py` CLI (reference
optimize_whole_sequence.py:5-117): same flags and defaults, same iteration over the
natural-sorted sub-directories of --data_path, same printed summary, running on the B200 path.

Added flags (not in the reference): --max_iter, --seed, --local_vae, --global_vae, --batch_clips (default True: all
clips of the dataset are optimised in one batched solve; False = one `main` call per clip, as the reference loops).
"""
import argparse
import os
import re

import numpy as np


def natsorted(names):
    """Natural sort (the reference uses natsort.natsorted, optimize_whole_sequence.py:48)."""
    def key(s):
        return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)]
    return sorted(names, key=key)


SUMMARY = [
    ("Average original global pose mpjpe", "original_global_mpjpe"),
    ("Average mid global pose mpjpe", "mid_global_mpjpe"),
    ("Average optimized global pose mpjpe", "optimized_global_mpjpe"),
    None,
    ("Average original cam pose error", "original_camera_pos_error"),
    ("Average optimized cam pose error", "optimized_camera_pos_error"),
    None,
    ("Average original aligned cam pose error", "original_aligned_camera_pos_error"),
    ("Average optimized aligned cam pose error", "optimized_aligned_camera_pos_error"),
    None,
    ("Average original_aligned_global_mpjpe", "original_aligned_global_mpjpe"),
    ("Average aligned_mid_seq_mpjpe", "aligned_mid_seq_mpjpe"),
    ("Average optimized_aligned_global_mpjpe", "optimized_aligned_global_mpjpe"),
    None,
    ("Average aligned original global pose mpjpe", "aligned_original_mpjpe"),
    ("Average aligned mid local pose mpjpe", "aligned_mid_optimized_mpjpe"),
    ("Average aligned optimized global pose mpjpe", "aligned_optimized_mpjpe"),
    None,
    ("Average bone length aligned original global pose mpjpe", "bone_length_aligned_original_mpjpe"),
    ("Average bone length aligned mid local pose mpjpe", "bone_length_aligned_mid_optimized_mpjpe"),
    ("Average bone length aligned optimized global pose mpjpe", "bone_length_aligned_optimized_mpjpe"),
    None,
]


def run(args):
    from globalegomocap_b200 import optimizer as gem
    if args.seed is not None:
        import torch
        torch.manual_seed(args.seed)
    collected = {k[1]: [] for k in SUMMARY if k}
    joints_error = []
    kwargs = dict(camera_model_path=args.camera, vae_weight=args.vae, gmm_weight=args.gmm, smoothness_weight=args.smooth,
                  visualization=False, save=args.save, bone_length_weight=args.bone_length, weight_3d=args.weight_3d,
                  reproj_weight=args.reproj_weight, merge=args.merge, final_smooth=args.final_smooth,
                  max_iter=args.max_iter, local_vae_path=args.local_vae, global_vae_path=args.global_vae)
    data_paths = []
    for name in natsorted(os.listdir(args.data_path)):
        data_path = os.path.join(args.data_path, name)
        print("running data: {}".format(data_path))
        if not os.path.isdir(data_path):
            continue
        data_paths.append(data_path)
    if args.batch_clips:
        # every window of every clip in one batched solve (a 100-frame clip alone is 12 windows: far too few to fill a GPU)
        results = gem.main_batch(data_paths, **kwargs)
    else:
        results = [gem.main(p, **kwargs) for p in data_paths]
    for res, _est, _mid, _opt, _gt in results:
        for k in collected:
            collected[k].append(res[k])
        joints_error.append(res["joints_error"])
        if res["bone_length_aligned_optimized_mpjpe"] > res["bone_length_aligned_mid_optimized_mpjpe"]:
            print(res)
    for item in SUMMARY:
        if item is None:
            print("-----------------------------------------")
        else:
            print("{}: {}".format(item[0], np.average(collected[item[1]])))
    print("joints error is: {}".format(np.mean(joints_error, axis=0)))
    print("-------------------------------------------------------------")


if __name__ == "__main__":
    from globalegomocap_b200.optimizer import GLOBAL_VAE_PATH, LOCAL_VAE_PATH

    def boolean(x):
        return str(x).lower() == "true"

    parser = argparse.ArgumentParser(description="Data directory number")
    parser.add_argument("--data_path", required=True, type=str)
    parser.add_argument("--camera", required=False, type=str, default="utils/fisheye/fisheye.calibration.json")
    parser.add_argument("--vae", required=False, type=float, default=0.00)
    parser.add_argument("--gmm", required=False, type=float, default=0.00)
    parser.add_argument("--smooth", required=False, type=float, default=0.001)
    parser.add_argument("--bone_length", required=False, type=float, default=0.01)
    parser.add_argument("--weight_3d", required=False, type=float, default=0.01)
    parser.add_argument("--reproj_weight", required=False, type=float, default=0.01)
    parser.add_argument("--save", required=False, default=False, type=boolean)
    parser.add_argument("--final_smooth", required=False, default=True, type=boolean)
    parser.add_argument("--merge", required=False, default=True, type=boolean)
    parser.add_argument("--max_iter", required=False, default=25, type=int)
    parser.add_argument("--seed", required=False, default=None, type=int)
    parser.add_argument("--batch_clips", required=False, default=True, type=boolean)
    parser.add_argument("--local_vae", required=False, default=LOCAL_VAE_PATH, type=str)
    parser.add_argument("--global_vae", required=False, default=GLOBAL_VAE_PATH, type=str)
    run(parser.parse_args())
