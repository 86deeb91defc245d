"""Initial 3-D lift: heat maps + per-joint depths -> local skeletons.

Host-side mirror of the reference's `Skeleton.set_skeleton` / `set_skeleton_from_file` (utils/skeleton.py:33-46,
73-88) for whole sequences at once: the argmax over the maps and the fisheye back-projection run in ONE CUDA kernel
(`gem_lift_skeleton`, csrc/lift.cu) that streams the maps in their pickle layout; the optional bone-length
normalisation (`_skeleton_resize`, :123-135) is sequential along the kinematic chain and tiny, and stays on the host
in float64.  There is no CPU fallback: without the CUDA library the call raises GemError.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np
import torch

from ._lib import GemError, check, load

KINEMATIC_PARENTS = [0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13]      # utils/skeleton.py:22


def load_camera_c2w(path):
    """(polynomialC2W float64[P], cx, cy) from a calibration JSON with the reference's keys
    (FishEyeCalibrated.py:8-14)."""
    with open(path) as f:
        cal = json.load(f)
    intr = np.asarray(cal["intrinsic"], dtype=np.float64)
    return np.asarray(cal["polynomialC2W"], dtype=np.float64), float(intr[0][2]), float(intr[1][2])


def _tree_levels(parents):
    """Joints grouped by their depth in the kinematic tree (the root, which is its own parent, is depth 0)."""
    parents = list(parents)
    depth = [0] * len(parents)
    for j, p in enumerate(parents):            # parents precede their children in the reference's joint order
        depth[j] = 0 if p == j else depth[p] + 1
    return [np.asarray([j for j, d in enumerate(depth) if d == lvl]) for lvl in range(1, max(depth) + 1)]


def rescale_bones(joints, bone_length_mm, parents=KINEMATIC_PARENTS):
    """Every bone of every pose rescaled to a given length, the tree re-grown from the root
    (`Skeleton._skeleton_resize`, utils/skeleton.py:123-135): joints [..., J, 3] float64 in metres, bone_length_mm [J]
    in millimetres (entry 0 unused).  One vectorised update per tree depth, over all poses at once; a joint is its
    parent's NEW position plus its own rescaled bone, added in the reference's order, so the result is the
    reference's to the last bit."""
    pts = np.array(joints, dtype=np.float64)
    par = np.asarray(parents)
    bone = pts - pts[..., par, :]
    length = np.linalg.norm(bone, axis=-1)
    target = np.asarray(bone_length_mm, dtype=np.float64)
    scale = np.zeros_like(length)
    scale[..., 1:] = target[1:] / length[..., 1:]
    bone = bone * scale[..., None] / 1000
    for level in _tree_levels(parents):
        pts[..., level, :] = pts[..., par[level], :] + bone[..., level, :]
    return pts


def skeleton_resize(points_3d, bone_length):
    """`Skeleton._skeleton_resize` for one pose or a stack of poses (see rescale_bones)."""
    return rescale_bones(points_3d, bone_length)


def lift_skeletons(heatmaps, depths, poly_c2w, cx, cy, bone_length=None, device=None, upscale=16, pad_x=128,
                   return_preds=False):
    """heatmaps [N][H][W][J] float32 (HWC, as in the pickles; numpy or torch, host or device), depths [N][J] ->
    local skeletons [N][J][3] float64 (numpy).  With `return_preds`: also (preds [N][J][2], maxvals [N][J])."""
    if not torch.cuda.is_available():
        raise GemError("no CUDA device: the lift has no CPU fallback")
    lib = load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    heat = torch.as_tensor(heatmaps).to(device=dev, dtype=torch.float32).contiguous()
    depth = torch.as_tensor(np.asarray(depths, dtype=np.float64) if not torch.is_tensor(depths) else depths)
    depth = depth.to(device=dev, dtype=torch.float64).contiguous()
    n, h, w, j = heat.shape
    if tuple(depth.shape) != (n, j):
        raise ValueError("depths must be [N][J]")
    points = torch.empty((n, j, 3), dtype=torch.float64, device=dev)
    preds = torch.empty((n, j, 2), dtype=torch.float32, device=dev)
    maxvals = torch.empty((n, j), dtype=torch.float32, device=dev)
    poly = np.ascontiguousarray(poly_c2w, dtype=np.float64)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib.gem_lift_skeleton(stream, n, h, w, j, C.c_void_p(heat.data_ptr()), C.c_void_p(depth.data_ptr()),
                                    poly.ctypes.data_as(C.POINTER(C.c_double)), int(poly.size), float(cx), float(cy),
                                    int(upscale), int(pad_x), C.c_void_p(points.data_ptr()),
                                    C.c_void_p(preds.data_ptr()), C.c_void_p(maxvals.data_ptr()), None))
    out = points.cpu().numpy()
    if bone_length is not None:
        out = rescale_bones(out, bone_length)
    if return_preds:
        return out, preds.cpu().numpy(), maxvals.cpu().numpy()
    return out
