"""Error metrics returned by ``main`` (reference calculate_errors.py:114-179):
MPJPE variants with no / per-sequence / per-pose similarity alignment and with
bone-length normalisation (SURVEY.md §8f row N2 — not part of the optimisation hot path).

The per-pose alignments — one 3x3 SVD per frame and variant, six Python loops over the frames in the
reference — run on the GPU (`gem_pose_align_errors`, csrc/metrics.cu) when ``on_device=True`` (what ``main``
uses); the handful of whole-sequence reductions stay in numpy.  ``on_device=False`` is the same arithmetic in
numpy for callers that only hold host arrays.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

PARENTS = [0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13]

# mean skeleton in millimetres, joints x xyz (the `mean3D` array of the reference's
# utils/fisheye/mean3D.mat, transposed): only its bone lengths are used.
MEAN3D_MM = np.array([
    [6.12454847, 233.90813433, 176.25176082], [145.97761, 232.60823975, 220.73112637],
    [258.72083056, 188.18493809, 404.39836013], [281.27554815, 72.79136312, 488.37987609],
    [-130.58758154, 239.16565076, 232.02432922], [-217.63663461, 203.68825151, 436.14841643],
    [-234.47818229, 91.05888921, 529.22255096], [122.57391072, 239.95855861, 675.05067301],
    [157.99031993, 133.01398165, 1019.17833662], [172.09879492, 176.20098748, 1331.949378],
    [215.33356937, 37.42165039, 1391.75072893], [-52.15750419, 243.04617535, 683.67509016],
    [-59.0959752, 149.38252591, 1037.58363271], [-36.18717374, 180.44482382, 1353.00767289],
    [-80.10264932, 44.79721165, 1407.87463384]])


def umeyama(P, Q):
    """Similarity (c, R, t) minimising sum |c P R + t - Q|^2 (utils/rigid_transform_with_scale.py:18-43)."""
    n = P.shape[0]
    cp, cq = P - P.mean(axis=0), Q - Q.mean(axis=0)
    C = cp.T.dot(cq) / n
    V, S, Wt = np.linalg.svd(C)
    if np.linalg.det(V) * np.linalg.det(Wt) < 0.0:
        S[-1] = -S[-1]
        V[:, -1] = -V[:, -1]
    R = V.dot(Wt)
    c = np.sum(S) / np.var(P, axis=0).sum()
    t = Q.mean(axis=0) - P.mean(axis=0).dot(c * R)
    return c, R, t


def _mpjpe(a, b):
    return np.mean(np.linalg.norm(np.asarray(a) - np.asarray(b), axis=-1))


def _root_error(est, gt):
    est, gt = np.asarray(est), np.asarray(gt)
    return np.mean(np.linalg.norm((est[:, 7] + est[:, 11]) / 2 - (gt[:, 7] + gt[:, 11]) / 2, axis=1))


def _align_sequence(est, gt):
    p, q = np.asarray(est).reshape(-1, 3), np.asarray(gt).reshape(-1, 3)
    c, R, t = umeyama(p, q)
    return (p.dot(R) * c + t).reshape(-1, 15, 3)


def resize_to_mean_bones(joints):
    """Poses [..., 15, 3] with every bone rescaled to the mean skeleton's length (`Skeleton.skeleton_resize_single`,
    utils/skeleton.py:118-122 with mean3D.mat's lengths)."""
    from .lift import rescale_bones
    return rescale_bones(joints, np.linalg.norm(MEAN3D_MM - MEAN3D_MM[PARENTS], axis=1), PARENTS)


def _align_per_pose_host(est, gt, resize):
    est, gt = np.array(est, dtype=np.float64), np.array(gt, dtype=np.float64)
    if resize:
        est, gt = resize_to_mean_bones(est), resize_to_mean_bones(gt)
    out = np.zeros_like(est)
    for i in range(est.shape[0]):
        c, R, t = umeyama(est[i], gt[i])
        out[i] = est[i].dot(R) * c + t
    return out, gt


def _align_per_pose_device(est, gt, resize):
    """`_align_per_pose` on the GPU: one thread per frame (csrc/metrics.cu); raises GemError without the CUDA library."""
    import ctypes as C

    import torch

    from ._lib import GemError, check, load
    if not torch.cuda.is_available():
        raise GemError("no CUDA device: use calculate_errors(..., on_device=False) for host arrays")
    lib = load()
    dev = torch.device("cuda", torch.cuda.current_device())
    e = torch.as_tensor(np.ascontiguousarray(est, dtype=np.float64), device=dev)
    g = torch.as_tensor(np.ascontiguousarray(gt, dtype=np.float64), device=dev)
    n, j = e.shape[0], e.shape[1]
    aligned, gt_out = torch.empty_like(e), torch.empty_like(g)
    err = torch.empty((n, j), dtype=torch.float64, device=dev)
    parents = (C.c_int32 * j)(*PARENTS[:j])
    bone = None
    if resize:
        bl = np.linalg.norm(MEAN3D_MM - MEAN3D_MM[PARENTS], axis=1)
        bone = (C.c_double * j)(*bl[:j])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.gem_pose_align_errors(stream, n, j, C.c_void_p(e.data_ptr()), C.c_void_p(g.data_ptr()), parents, bone,
                                    C.c_void_p(aligned.data_ptr()), C.c_void_p(gt_out.data_ptr()),
                                    C.c_void_p(err.data_ptr())))
    return aligned.cpu().numpy(), gt_out.cpu().numpy()


def calculate_errors(final_estimated_seq, mid_estimated_seq, final_optimized_seq, final_gt_seq, on_device=False):
    est, mid, opt, gt = (np.asarray(a, dtype=np.float64) for a in
                         (final_estimated_seq, mid_estimated_seq, final_optimized_seq, final_gt_seq))
    _align_per_pose = _align_per_pose_device if on_device else _align_per_pose_host
    res = OrderedDict()
    res["original_global_mpjpe"] = _mpjpe(est, gt)
    res["mid_global_mpjpe"] = _mpjpe(mid, gt)
    res["optimized_global_mpjpe"] = _mpjpe(opt, gt)
    res["original_camera_pos_error"] = _root_error(est, gt)
    res["optimized_camera_pos_error"] = _root_error(opt, gt)
    a_est, a_mid, a_opt = _align_sequence(est, gt), _align_sequence(mid, gt), _align_sequence(opt, gt)
    res["original_aligned_camera_pos_error"] = _root_error(a_est, gt)
    res["mid_aligned_camera_pose_error"] = _root_error(a_mid, gt)
    res["optimized_aligned_camera_pos_error"] = _root_error(a_opt, gt)
    res["original_aligned_global_mpjpe"] = _mpjpe(a_est, gt)
    res["aligned_mid_seq_mpjpe"] = _mpjpe(a_mid, gt)
    res["optimized_aligned_global_mpjpe"] = _mpjpe(a_opt, gt)
    p_est, _ = _align_per_pose(est, gt, False)
    p_mid, _ = _align_per_pose(mid, gt, False)
    p_opt, _ = _align_per_pose(opt, gt, False)
    res["aligned_original_mpjpe"] = _mpjpe(p_est, gt)
    res["aligned_mid_optimized_mpjpe"] = _mpjpe(p_mid, gt)
    res["aligned_optimized_mpjpe"] = _mpjpe(p_opt, gt)
    # the reference re-assigns the (resized) ground truth between the three calls
    # (calculate_errors.py:145-147); resizing is idempotent, so one resize is equivalent
    b_est, gt_r = _align_per_pose(est, gt, True)
    b_mid, gt_r = _align_per_pose(mid, gt_r, True)
    b_opt, gt_r = _align_per_pose(opt, gt_r, True)
    res["bone_length_aligned_original_mpjpe"] = _mpjpe(b_est, gt_r)
    res["bone_length_aligned_mid_optimized_mpjpe"] = _mpjpe(b_mid, gt_r)
    res["bone_length_aligned_optimized_mpjpe"] = _mpjpe(b_opt, gt_r)
    res["joints_error"] = np.mean(np.linalg.norm(b_opt - gt_r, axis=2), axis=0)
    return res
