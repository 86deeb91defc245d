"""numpy restatement of the reference's initial 3-D lift: heat-map argmax + per-joint depth ->
local skeleton, the step that produces the optimiser's `estimated_local_skeleton` input.

Follows reference utils/skeleton.py:33-46 (`Skeleton.set_skeleton`), :73-88 (`set_skeleton_from_file`: nearest
resize of the 64x64 maps to 1024x1024 and 128-pixel padding left and right), :176-204 (`get_max_preds`),
:123-135 (`_skeleton_resize`) and utils/fisheye/FishEyeCalibrated.py:18-33 (`camera2world`, float64).

Oracle: test infrastructure only (see oracle/__init__.py).  Pinned by tests/test_oracle_lift.py against
tests/golden/lift.npz, produced by tests/golden/make_golden_lift.py from the unmodified reference.
"""
from __future__ import annotations

import numpy as np

KINEMATIC_PARENTS = [0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13]     # utils/skeleton.py:22
UPSCALE = 16           # cv2.resize(64 -> 1024, INTER_NEAREST): destination pixel i shows source texel i // 16
PAD_X = 128            # np.pad(((0, 0), (128, 128), (0, 0))): 1024 -> 1280 columns


def max_preds(heat_hwc):
    """`get_max_preds` (utils/skeleton.py:176-204) applied to the resized + padded maps of
    `set_skeleton_from_file` (:80-82), evaluated on the 64x64 source maps.

    heat_hwc: [N][H][W][J] float32 (the pickle's layout).  Returns (preds [N][J][2] float32 image coordinates
    (x, y) in the 1280x1024 frame, maxvals [N][J] float32, argmax [N][J] int32 = y0 * W + x0 of the first maximum
    in row-major order).  The first maximum of the nearest-resized, padded image is the top-left pixel of the block of
    the source map's first row-major maximum: (16 x0 + 128, 16 y0).  Maps whose maximum is <= 0 give (0, 0)
    (`pred_mask`, :199-203); a map that is <= 0 everywhere has its maximum, 0, in the padding: pixel (0, 0).
    """
    heat = np.asarray(heat_hwc)
    n, h, w, j = heat.shape
    flat = heat.transpose(0, 3, 1, 2).reshape(n, j, h * w)
    idx = np.argmax(flat, axis=2)
    maxvals = np.take_along_axis(flat, idx[..., None], axis=2)[..., 0]
    x0 = (idx % w).astype(np.float32)
    y0 = (idx // w).astype(np.float32)
    preds = np.stack([x0 * UPSCALE + PAD_X, y0 * UPSCALE], axis=-1).astype(np.float32)
    # the padded image holds zeros left and right: a map with no positive texel has its first maximum (0) at pixel
    # 0 of row 0 when max < 0, or wherever the first zero is when max == 0 — either way the mask zeroes the result
    mask = (maxvals > 0.0).astype(np.float32)[..., None]
    preds = preds * mask
    maxvals = np.where(maxvals > 0.0, maxvals, np.maximum(maxvals, 0.0)).astype(np.float32)   # padding zeros win
    return preds, maxvals, idx.astype(np.int32)


def camera2world(points, depth, center, poly_c2w):
    """`FishEyeCameraCalibrated.camera2world` (FishEyeCalibrated.py:18-33), float64: points [n][2], depth [n]."""
    p = np.asarray(points, dtype=np.float64) - np.asarray(center, dtype=np.float64)
    x, y = p[:, 0], p[:, 1]
    dist = np.sqrt(np.square(x) + np.square(y))
    z = np.polyval(np.asarray(poly_c2w, dtype=np.float64)[::-1], dist)
    v = np.array([x, y, -z])
    norm = np.linalg.norm(v, axis=0)
    return (v / norm * np.asarray(depth, dtype=np.float64)).transpose()


def skeleton_resize(points_3d, bone_length):
    """`_skeleton_resize` (utils/skeleton.py:123-135): bones rescaled to `bone_length` (in mm) along the kinematic
    chain, in place order of the joints as the reference loops."""
    pts = np.array(points_3d, dtype=np.float64)
    vec = pts - pts[KINEMATIC_PARENTS, :]
    est = np.linalg.norm(vec, axis=1)
    multi = np.concatenate(([0.0], np.asarray(bone_length, dtype=np.float64)[1:] / est[1:]))
    resized = vec * multi[:, None] / 1000
    out = pts
    for i in range(out.shape[0]):
        out[i, :] = out[KINEMATIC_PARENTS[i], :] + resized[i, :]
    return out


def lift(heat_hwc, depth, center, poly_c2w, bone_length=None):
    """`set_skeleton` for every frame: [N][H][W][J] maps + [N][J] depths -> [N][J][3] float64 local skeletons."""
    preds, maxvals, idx = max_preds(heat_hwc)
    out = np.stack([camera2world(preds[i], depth[i], center, poly_c2w) for i in range(preds.shape[0])])
    if bone_length is not None:
        out = np.stack([skeleton_resize(o, bone_length) for o in out])
    return out, preds, maxvals, idx
