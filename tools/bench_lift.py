#!/usr/bin/env python
"""Throughput of the initial 3-D lift kernel (gem_lift_skeleton): frames/s and HBM GB/s against the measured peak.

    python tools/bench_lift.py [--frames 15000] [--steps 10]

Algorithmic bytes per frame = the frame's maps, read once: 64*64*15*4 = 245,760 B (+ 0.5 KB of depths / results).
Inputs (3.7 GB at the default size) are larger than L2, so no flush is needed between launches.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=15000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    from globalegomocap_b200._lib import check, load
    from globalegomocap_b200.lift import load_camera_c2w
    from globalegomocap_b200.synthetic import DEFAULT_CAMERA_JSON
    lib = load()
    poly, cx, cy = load_camera_c2w(DEFAULT_CAMERA_JSON)
    dev = torch.device("cuda", 0)
    n, h, w, j = args.frames, 64, 64, 15
    g = torch.Generator(device=dev).manual_seed(0)
    heat = torch.rand((n, h, w, j), dtype=torch.float32, device=dev, generator=g)
    depth = torch.rand((n, j), dtype=torch.float64, device=dev, generator=g) + 0.5
    points = torch.empty((n, j, 3), dtype=torch.float64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run():
        check(lib.gem_lift_skeleton(stream, n, h, w, j, C.c_void_p(heat.data_ptr()), C.c_void_p(depth.data_ptr()),
                                    poly.ctypes.data_as(C.POINTER(C.c_double)), int(poly.size), cx, cy, 16, 128,
                                    C.c_void_p(points.data_ptr()), None, None, None))

    for _ in range(args.warmup):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    bytes_per_launch = n * (h * w * j * 4 + j * 8 + j * 3 * 8)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gbs = bytes_per_launch / (ms / 1e3) / 1e9
    print(json.dumps({"kernel": "lift_kernel (gem_lift_skeleton)", "frames": n, "ms_per_launch": ms,
                      "frames_per_s": n / (ms / 1e3), "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak,
                                                                   "unit": "GB/s", "frac": gbs / peak,
                                                                   "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"},
                      "algorithmic_bytes_per_frame": h * w * j * 4 + j * 8 + j * 24}))


if __name__ == "__main__":
    main()
