#!/usr/bin/env python
"""Generates tests/golden/dist.npz and tests/golden/energy_cam14.npz by running the UNMODIFIED
reference (/root/reference) in this CPU container (same harness as make_golden.py).

dist.npz — the evidence behind "free-running parity is statistical for the reference too"
(reference optimizer.py:261-270):

  * 64 windows (one synthetic 514-frame clip, regenerated from its seed by the tests) x 2 stages x
    max_iter in {3, 25}, solved by the reference with torch at ONE thread and again at N threads
    (only the GEMM reduction order differs).  Both runs of the global stage are anchored at the
    1-thread local result, so every stage's two runs see identical inputs.
  * per run: the per-evaluation energies, n_iter / func_evals, the final window pose;
  * for the 1-thread runs also every `_cubic_interpolate` call of torch's strong-Wolfe search
    (torch/optim/lbfgs.py:12-37): which evaluation it produced, its arguments, the discriminant
    d1^2 - g1*g2 and the step it returns when the newer loss value moves by +-1 / +-2 fp32 ulps —
    the measured amplification of one ulp of the loss into the step length.

The `-m gpu` test compares CUDA-vs-reference(1 thread) deviations with reference(1 thread)-vs-
reference(N threads) deviations on the same windows, quantile by quantile.

Usage (container only):  python tests/golden/make_golden_dist.py [--windows 64] [--threads 8]
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
import make_golden as mg  # noqa: E402  (harness shared with the other goldens)

from globalegomocap_b200 import synthetic as syn  # noqa: E402

CLIP_SEED = 31
MAX_EVALS = 34          # columns of the energy table (max_eval + 1 = 32 at max_iter 25)


def f32_step(v, k):
    """v moved by k fp32 ulps (v is a python float holding an fp32 value)."""
    a = np.float32(v)
    for _ in range(abs(k)):
        a = np.nextafter(a, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
    return float(a)


class CubicTracer:
    """Records every _cubic_interpolate call of torch's strong-Wolfe search."""

    def __init__(self, counter):
        import torch.optim.lbfgs as tl
        self.tl = tl
        self.orig = tl._cubic_interpolate
        self.rows = []
        self.counter = counter          # callable: closure evaluations done so far in this solve
        tr = self

        def traced(x1, f1, g1, x2, f2, g2, bounds=None):
            res = tr.orig(x1, f1, g1, x2, f2, g2, bounds)
            fl = lambda v: float(v)     # noqa: E731
            d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
            d2s = d1 ** 2 - g1 * g2
            pert = []
            for k in (-2, -1, 1, 2):
                try:
                    pert.append(fl(tr.orig(x1, f1, g1, x2, f32_step(fl(f2), k), g2, bounds)))
                except Exception:
                    pert.append(float("nan"))
            lo, hi = (fl(bounds[0]), fl(bounds[1])) if bounds is not None else (float("nan"), float("nan"))
            tr.rows.append([float(tr.counter()), fl(x1), fl(f1), fl(g1), fl(x2), fl(f2), fl(g2), fl(d1), fl(d2s),
                            fl(res), lo, hi] + pert)
            return res

        tl._cubic_interpolate = traced

    def take(self):
        r, self.rows = self.rows, []
        return np.asarray(r, dtype=np.float64).reshape(-1, 16)

    def restore(self):
        self.tl._cubic_interpolate = self.orig


def gen_dist(h: mg.Harness, n_windows: int, threads: int, max_iters=(3, 25)):
    import torch
    inj = mg.EpsInjector(h.ref)
    tracer = mg.LossTracer(h.ref)
    cubic = CubicTracer(lambda: len(tracer.records))
    est = np.asarray(h.clip["estimated_local_skeleton"])
    heat_all = np.asarray(h.clip["heatmap_list"])
    cams = np.asarray(h.clip["camera_pose_list"])
    starts = syn.window_starts(len(est))[:n_windows]
    W = len(starts)
    rng = np.random.default_rng(9001)
    eps = rng.standard_normal((W, 2, 2048)).astype(np.float32)
    out = {"clip_seed": np.int64(CLIP_SEED), "n_frames": np.int64(len(est)), "starts": np.asarray(starts, np.int64),
           "eps": eps, "threads": np.asarray([1, threads], np.int64), "max_iters": np.asarray(max_iters, np.int64),
           "heat_checksum": np.float64(heat_all.astype(np.float64).sum()),
           "cubic_columns": np.asarray(["eval_index", "x1", "f1", "g1", "x2", "f2", "g2", "d1", "d2_square", "t",
                                        "lo", "hi", "t_m2ulp", "t_m1ulp", "t_p1ulp", "t_p2ulp"])}
    try:
        for mi in max_iters:
            E = np.full((2, 2, W, MAX_EVALS), np.nan)              # [thread setting][stage][window][evaluation]
            pose = np.zeros((2, 2, W, 10, 15, 3), np.float32)
            n_iter = np.zeros((2, 2, W), np.int32)
            n_eval = np.zeros((2, 2, W), np.int32)
            rel_in = np.zeros((W, 10, 15, 3), np.float64)           # the global stage's anchor (1-thread local result)
            cub_rows, cub_index = [], []
            t0 = time.time()
            for ti, nthr in enumerate((1, threads)):
                torch.set_num_threads(nthr)
                lopt = h.make_optimizer("local", mg.W_LOCAL, max_iter=mi)
                gopt = h.make_optimizer("global", mg.W_GLOBAL, max_iter=mi)
                for wi, s in enumerate(starts):
                    for si, stage in enumerate(("local", "global")):
                        inj.push(eps[wi, si])
                        if stage == "local":
                            x0 = est[s:s + 10]
                            res = lopt.optimize_pose_seq_pytorch_LBFGS(x0, heat_all[s:s + 10], x0.copy())
                            if ti == 0:
                                rel_in[wi] = h.ref.get_relative_global_pose_with_camera_matrix(res, cams[s:s + 10])
                            opt_obj = lopt
                        else:
                            x0 = rel_in[wi]
                            res = gopt.optimize_pose_seq_pytorch_LBFGS(x0, heat_all[s:s + 10], x0.copy())
                            opt_obj = gopt
                        rec = tracer.take()
                        rows = cubic.take()
                        E[ti, si, wi, :len(rec)] = [r[0] for r in rec]
                        pose[ti, si, wi] = res
                        n_eval[ti, si, wi] = len(rec)
                        if ti == 0:
                            cub_index.append([si, wi, len(rows)])
                            cub_rows.append(rows)
                print(f"max_iter={mi} threads={nthr}: {time.time() - t0:.0f} s", flush=True)
            out[f"mi{mi}_E"] = E
            out[f"mi{mi}_pose"] = pose
            out[f"mi{mi}_n_eval"] = n_eval
            out[f"mi{mi}_rel_in"] = rel_in
            out[f"mi{mi}_cubic"] = np.concatenate(cub_rows, axis=0)
            out[f"mi{mi}_cubic_index"] = np.asarray(cub_index, np.int64)      # (stage, window, rows) in storage order
        out["mean_bone_length"] = lopt.mean_bone_length.numpy()
    finally:
        cubic.restore()
        tracer.restore()
        inj.restore()
        torch.set_num_threads(1)
    np.savez_compressed(os.path.join(mg.OUT, "dist.npz"), **out)
    print("dist:", {k: np.asarray(v).shape for k, v in out.items()})


def gen_energy_cam14():
    """Energy terms + autograd gradients with the 14-coefficient calibration
    (utils/fisheye/pose_fisheye_fisheye.calibration_new.json, FishEyeCalibrated.py:8-14).  The clip's own maps
    were drawn for the 11-coefficient camera, so every case samples the dense synthetic maps."""
    import torch
    os.chdir(mg.REPO)
    clip = syn.make_clip(58, seed=7)
    h = mg.Harness(tempfile.mkdtemp(prefix="gem_golden_c14_"), clip, cam="new")
    out, names = {}, []
    for wname, weights in (("local", mg.W_LOCAL), ("all", mg.W_ALL)):
        opt = h.make_optimizer("local", weights)
        for cname, start, xnp in mg.energy_cases(h):
            x0 = np.asarray(clip["estimated_local_skeleton"])[start:start + 10]
            heat = syn.dense_heat_window(2)
            opt.initial_pose = torch.from_numpy(x0).float()
            hs = torch.from_numpy(heat).float().permute((0, 3, 1, 2)).contiguous()
            opt.heatmap_seq = hs.view(-1, hs.shape[-2], hs.shape[-1])
            key = f"{wname}__{cname}"
            names.append(key)
            x = torch.from_numpy(xnp).float().requires_grad_(True)
            e = opt.reprojection_energy_heatmap_fast(x)
            out[f"{key}__E_reproj"] = np.float32(e.item())
            out[f"{key}__G_reproj"] = torch.autograd.grad(e, x)[0].numpy()
            x = torch.from_numpy(xnp).float().requires_grad_(True)
            tot = (opt.weight_3d * opt.pose_energy_3d(x) + opt.smooth_weight * opt.smooth_accelerate(x)
                   + opt.bone_length_weight * opt.bone_length_energy(x) + opt.vae_weight * opt.vae_energy(x)
                   + opt.reproj_weight * opt.reprojection_energy_heatmap_fast(x))
            out[f"{key}__E_total"] = np.float32(tot.item())
            out[f"{key}__G_total"] = torch.autograd.grad(tot, x)[0].numpy()
            out[f"{key}__x"] = xnp
            out[f"{key}__start"] = np.int64(start)
        out[f"{wname}__weights"] = np.asarray([weights["weight_3d"], weights["smooth_weight"],
                                               weights["bone_length_weight"], weights["vae_weight"],
                                               weights["reproj_weight"]], dtype=np.float64)
        out["mean_bone_length"] = opt.mean_bone_length.numpy()
    out["names"] = np.asarray(names)
    out["n_poly"] = np.int64(len(syn.load_camera(syn.DEFAULT_CAMERA_JSON.replace("calibration.json",
                                                                                  "calibration_new.json"))[0]))
    np.savez_compressed(os.path.join(mg.OUT, "energy_cam14.npz"), **out)
    print("energy_cam14:", len(names), "cases, n_poly", int(out["n_poly"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=64)
    ap.add_argument("--threads", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    todo = args.only or ["cam14", "dist"]
    if "cam14" in todo:
        gen_energy_cam14()
    if "dist" in todo:
        os.chdir(mg.REPO)
        clip = syn.make_clip(8 * (args.windows - 1) + 10, seed=CLIP_SEED)
        h = mg.Harness(tempfile.mkdtemp(prefix="gem_golden_dist_"), clip)
        gen_dist(h, args.windows, args.threads)


if __name__ == "__main__":
    main()
