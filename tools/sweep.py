#!/usr/bin/env python
"""Throughput sweep over the number of concurrent windows (BASELINE.json configs[4]: 1e3-1e5 windows).

    python tools/sweep.py --sequences 27 --frames 3000 --replicate 10      # ~1e5 windows on one GPU

Windows of `--sequences` synthetic sequences are optimised together; `--replicate R` runs every window R
times (the replicas share the sequences' heat maps in HBM, as SURVEY.md 8d suggests for the 1e5 point, but
start from their own reparameterisation noise, so they are independent solves).  Prints one JSON line:
whole-path frames/s (device-resident inputs, CUDA events) and, from a second single-stream pass with an
event pair around every launch, the per-kernel durations with the rates the rooflines are quoted in.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=3)
    ap.add_argument("--frames", type=int, default=3000)
    ap.add_argument("--replicate", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--chunks", type=int, default=-1)
    args = ap.parse_args()
    from globalegomocap_b200 import synthetic as syn
    from globalegomocap_b200.engine import Engine
    from globalegomocap_b200.pipeline import SequenceOptimizer, WindowBatch

    clips = [syn.make_clip(args.frames, seed=17 + s) for s in range(args.sequences)]
    bias = syn.mean_pose_bias(syn.make_clip(64, seed=17))
    weights = (syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias),
               syn.make_vae_state_dict(12, perturb_bn=True, pose_bias=bias))
    w_seq = sum(len(range(0, args.frames - 10 + 1, 8)) for _ in range(args.sequences))
    W = w_seq * args.replicate
    eng = Engine(max_windows=W)
    eng.set_camera(*syn.load_camera())
    eng.set_vae(0, weights[0]), eng.set_vae(1, weights[1])
    if args.chunks >= 1:
        eng.set_chunks(args.chunks)
    so = SequenceOptimizer(eng)
    batch = WindowBatch(eng, [{k: v for k, v in c.items() if k != "gt_global_skeleton"} for c in clips])
    if args.replicate > 1:
        R = args.replicate
        batch.frame_base, batch.clip_idx = batch.frame_base.repeat(R), batch.clip_idx.repeat(R)
        batch.frame_idx = batch.frame_idx.repeat(R, 1)
        batch.W = W
    torch.cuda.synchronize()

    def step():
        return so.solve(batch, eps=None)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        sol = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    evals = int(sol["local"]["func_evals"].sum().item() + sol["glob"]["func_evals"].sum().item())
    eng.read_profile()
    eng.set_profiling(True)
    step()
    prof = eng.read_profile()
    eng.set_profiling(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    per = {int(k): dict(launches=v[0], ms_avg=v[1] / max(v[0], 1), ms_total=v[1]) for k, v in prof.items()}
    out = {"windows": W, "sequences": args.sequences, "frames": args.frames, "replicate": args.replicate,
           "ms_per_step": ms, "frames_per_s": 8.0 * W / (ms / 1e3), "closure_evaluations": evals,
           "scratch_gb": eng.scratch_bytes() / 1e9}
    if 1 in per:
        e = per[1]
        byts = W * (7800 + 5400) / 2.0         # 7.8 KB (local) / 5.4 KB (global) per window per evaluation
        out["energy_kernel"] = {"ms_avg": e["ms_avg"], "GBps_algorithmic": byts / (e["ms_avg"] / 1e3) / 1e9,
                                "frac_of_measured_hbm": byts / (e["ms_avg"] / 1e3) / 1e9 / hbm}
    if 3 in per:
        e = per[3]
        # executed bytes of one advance: 10 vectors + 4(k-1) history rows of 8 KB per window that starts a new
        # iteration; over a stage k runs 1..24, on average ~54 rows -> 0.44 MB
        byts = W * 54 * 8192.0
        out["lbfgs_advance"] = {"ms_avg": e["ms_avg"], "GBps_estimated": byts / (e["ms_avg"] / 1e3) / 1e9,
                                "frac_of_measured_hbm": byts / (e["ms_avg"] / 1e3) / 1e9 / hbm}
    g_ms = sum(per[t]["ms_total"] for t in (100, 205) if t in per)
    g_n = sum(per[t]["launches"] for t in (100, 205) if t in per)
    if g_n:
        fl = 2.0 * W * 2048 * 2560
        out["latent_gemm"] = {"ms_avg": g_ms / g_n, "TFLOPs_fp32_equivalent": fl / (g_ms / g_n / 1e3) / 1e12,
                              "TFLOPs_tf32_executed": 3 * fl / (g_ms / g_n / 1e3) / 1e12,
                              "frac_of_measured_bf16": fl / (g_ms / g_n / 1e3) / 1e12 / tf}
    taps = [t for t in per if t in (101, 102, 103, 104, 105, 200, 201, 202, 203, 204)]
    if taps:
        fl = 2.0 * W * 10 * 2 * (768 * 128 + 384 * 64 + 192 * 64 * 2 + 192 * 45)
        t_ms = sum(per[t]["ms_total"] for t in taps) / max(per[taps[0]]["launches"], 1)
        out["conv_layers"] = {"ms_per_round": t_ms, "TFLOPs_fp32_equivalent": fl / (t_ms / 1e3) / 1e12}
    out["kernel_ms_total"] = {str(k): round(v["ms_total"], 3) for k, v in sorted(per.items())}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
