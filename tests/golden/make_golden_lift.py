#!/usr/bin/env python
"""Generates tests/golden/lift.npz by running the UNMODIFIED reference's initial 3-D lift
(/root/reference/utils/skeleton.py `Skeleton`, read-only) in this CPU container.

Each frame goes through the statements of `Skeleton.set_skeleton_from_file` (utils/skeleton.py:80-83: cv2 nearest
resize to 1024x1024, 128-pixel padding, transpose to CHW) and then the reference's own `get_max_preds`,
`set_skeleton` and `_skeleton_resize`.  Compatibility shims, none of which touch the arithmetic: `open3d` is stubbed
(mesh export only) and `numpy.float`, removed in numpy 1.24 and used at FishEyeCalibrated.py:23-24, is aliased to
`float`.

    python tests/golden/make_golden_lift.py
"""
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden", "lift.npz")


def main():
    for name in ("open3d", "natsort"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(np, "float"):
        np.float = float
    sys.path[:0] = [REF]
    import cv2
    from utils.skeleton import Skeleton
    cal = os.path.join(REF, "utils/fisheye/fisheye.calibration.json")
    sk = Skeleton(calibration_path=cal)

    rng = np.random.default_rng(20240607)
    n, h, w, j = 12, 64, 64, 15
    heat = rng.random((n, h, w, j), dtype=np.float32) * 0.2
    # Gaussian blobs like the pose network's maps
    yy, xx = np.mgrid[0:h, 0:w]
    for f in range(n):
        for k in range(j):
            cy, cx = rng.uniform(4, 60, 2)
            heat[f, :, :, k] += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 1.5 ** 2)).astype(np.float32)
    # edge cases: ties (the FIRST row-major maximum wins), maxima on the borders, an all-zero map, an all-negative
    # map, a map whose maximum is exactly zero
    heat[1, :, :, 0] = 0.0; heat[1, 5, 7, 0] = 3.0; heat[1, 5, 9, 0] = 3.0; heat[1, 40, 2, 0] = 3.0
    heat[1, :, :, 1] = 0.0; heat[1, 0, 0, 1] = 1.0
    heat[1, :, :, 2] = 0.0; heat[1, 63, 63, 2] = 1.0
    heat[1, :, :, 3] = 0.0
    heat[1, :, :, 4] = -1.0 - rng.random((h, w), dtype=np.float32)
    heat[1, :, :, 5] = -rng.random((h, w), dtype=np.float32); heat[1, 20, 30, 5] = 0.0
    heat[1, :, :, 6] = 0.0; heat[1, 0, 63, 6] = 2.0; heat[1, 63, 0, 6] = 2.0
    depth = rng.uniform(0.3, 2.5, (n, j))
    mean3d = rng.uniform(-400, 400, (j, 3))
    parents = Skeleton.kinematic_parents
    bone_length = np.linalg.norm(mean3d - mean3d[parents, :], axis=1)          # utils/skeleton.py:76-77

    preds, maxvals, points, resized = [], [], [], []
    for f in range(n):
        hm = cv2.resize(heat[f], dsize=(1024, 1024), interpolation=cv2.INTER_NEAREST)      # utils/skeleton.py:80
        hm = np.pad(hm, ((0, 0), (128, 128), (0, 0)), "constant", constant_values=0)        # :81
        hm = hm.transpose((2, 0, 1))                                                         # :82
        p, m = sk.get_max_preds(np.expand_dims(hm, axis=0))
        preds.append(p[0].copy()), maxvals.append(m[0, :, 0].copy())
        points.append(np.array(sk.set_skeleton(hm, depth[f], bone_length=None, to_mesh=False)))
        resized.append(np.array(sk.set_skeleton(hm, depth[f], bone_length=bone_length, to_mesh=False)))
    cam = sk.camera
    np.savez_compressed(OUT, heat=heat, depth=depth, bone_length=bone_length, preds=np.stack(preds),
                        maxvals=np.stack(maxvals), points=np.stack(points), resized=np.stack(resized),
                        center=np.asarray(cam.img_center, dtype=np.float64),
                        poly_c2w=np.asarray(cam.fisheye_polynomial, dtype=np.float64))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
