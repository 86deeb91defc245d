"""numpy restatement of the reference's per-sequence pipeline around the window
optimiser (reference optimizer.py:287-308, 311-450 and utils/utils.py:62-66,
99-112): window partition, SLAM camera transforms, the two L-BFGS stages,
overlap-average stitching and the final Gaussian smoothing.

Oracle: test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

from . import energy_np as en
from .lbfgs_np import lbfgs_minimize
from .vae_np import VaeNp

SEQ_LEN = 10
OVERLAP = 2


def window_starts(n_frames, seq_len=SEQ_LEN, overlap=OVERLAP):
    """optimizer.py:370 — trailing (N - seq_len) mod (seq_len - overlap) frames are dropped."""
    return list(range(0, n_frames - seq_len + 1, seq_len - overlap))


def transform_pose(pose, matrix):
    """utils/utils.py:62-66 (float64)."""
    homo = np.concatenate([pose, np.ones((pose.shape[0], 1))], axis=1)
    return matrix.dot(homo.T).T[:, :3]


def relative_global_pose(local_pose_seq, cam_seq):
    """get_relative_global_pose_with_camera_matrix, utils/utils.py:99-112:
    x'_t = inv(C_0) C_t x_t."""
    inv0 = np.linalg.inv(np.array(cam_seq[0], dtype=np.float64))
    return np.asarray([transform_pose(np.asarray(p, dtype=np.float64), inv0.dot(c))
                       for p, c in zip(local_pose_seq, cam_seq)])


def to_global_pose(rel_pose_seq, cam_seq):
    """relative_global_pose_to_global_pose, optimizer.py:302-308: C_0 x."""
    c0 = np.asarray(cam_seq[0], dtype=np.float64)
    return np.asarray([transform_pose(np.asarray(p, dtype=np.float64), c0) for p in rel_pose_seq])


def merge_batches(windows, overlap=OVERLAP):
    """optimizer.py:425-437: (W,T,15,3) -> (8W+2,15,3)."""
    windows = np.asarray(windows)
    if overlap == 0:
        return np.concatenate(windows)
    out = list(windows[0][:-overlap])
    for i in range(len(windows) - 1):
        a, b = windows[i], windows[i + 1]
        out.extend((a[-overlap:] + b[:overlap]) / 2)
        out.extend(b[overlap:-overlap])
    out.extend(windows[-1][-overlap:])
    return np.asarray(out)


def gaussian_kernel1d(sigma=1.0, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d, order 0."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def gaussian_filter1d_reflect(a, sigma=1.0, axis=0):
    """scipy.ndimage.gaussian_filter1d(a, sigma, axis, mode='reflect') as called
    at optimizer.py:450 ('reflect' = half-sample symmetric: d c b a | a b c d | d c b a)."""
    a = np.moveaxis(np.asarray(a, dtype=np.float64), axis, 0)
    k = gaussian_kernel1d(sigma)
    r = len(k) // 2
    n = a.shape[0]
    idx = np.arange(-r, n + r)
    period = 2 * n
    idx = np.mod(idx, period)
    idx = np.where(idx >= n, period - 1 - idx, idx)
    ext = a[idx]
    out = np.zeros_like(a)
    for j in range(2 * r + 1):
        out += k[j] * ext[j:j + n]
    return np.moveaxis(out, 0, axis)


class StageSolver:
    """One BodyPoseOptimizer (optimizer.py:33-276): a VAE, a camera, weights and
    the clip's mean bone lengths; ``solve`` is optimize_pose_seq_pytorch_LBFGS."""

    def __init__(self, state_dict, camera, mean_bone, weights, max_iter=25, lr=2):
        self.vae = VaeNp(state_dict, np.float32)
        self.poly, self.cx, self.cy = camera
        self.mean_bone = np.asarray(mean_bone, dtype=np.float32)
        self.weights = weights                      # (w3d, ws, wb, wv, wr)
        self.max_iter = max_iter
        self.lr = lr

    def closure_for(self, x0, heat):
        x0 = np.asarray(x0).astype(np.float32)

        def closure(z):
            pose, saved = self.vae.decode(z[None, :], keep=True)
            E, G, _ = en.total_energy(pose[0], x0, heat, self.mean_bone, self.weights, self.poly, self.cx, self.cy)
            gz = self.vae.decode_vjp(saved, G[None])
            return float(E), gz[0]

        return closure

    def solve(self, pose_in, heat, eps):
        x0 = np.asarray(pose_in).astype(np.float32)                     # optimizer.py:243
        z0 = self.vae.latent(x0.reshape(1, SEQ_LEN, 45), np.asarray(eps).reshape(1, -1))[0]
        z, info = lbfgs_minimize(self.closure_for(x0, heat), z0, lr=self.lr, max_iter=self.max_iter)
        return self.vae.decode(z[None, :])[0], info                     # optimizer.py:273-276


def run_sequence(clip, sd_local, sd_global, camera, eps, *, max_iter=25, final_smooth=True,
                 vae_weight=0.0, smooth=0.001, bone_length=0.01, weight_3d=0.01, reproj_weight=0.01):
    """optimizer.main (optimizer.py:311-450) without metrics/IO.  eps (W,2,latent):
    noise for local/global stage of each window in the reference's call order."""
    est = np.asarray(clip["estimated_local_skeleton"], dtype=np.float64)
    cams = np.asarray(clip["camera_pose_list"], dtype=np.float64)
    gt = np.asarray(clip["gt_global_skeleton"], dtype=np.float64)
    heat = np.asarray(clip["heatmap_list"])
    mb = en.mean_bone_length(est)
    # optimizer.py:352-358
    w_global = (weight_3d, smooth, 0.01, vae_weight, 0)
    w_local = (weight_3d / 10000, smooth / 100, bone_length, vae_weight, reproj_weight)
    local = StageSolver(sd_local, camera, mb, w_local, max_iter)
    glob = StageSolver(sd_global, camera, mb, w_global, max_iter)
    acc = {k: [] for k in ("est_global", "mid_local", "mid_global", "opt_global", "gt", "est_local")}
    infos = []
    for wi, s in enumerate(window_starts(len(est))):
        x_loc, c, h = est[s:s + SEQ_LEN], cams[s:s + SEQ_LEN], heat[s:s + SEQ_LEN]
        acc["est_local"].append(x_loc.copy())
        res_l, info_l = local.solve(x_loc, h, eps[wi, 0])
        acc["mid_local"].append(res_l)
        est_rel = relative_global_pose(x_loc, c)
        opt_rel = relative_global_pose(res_l, c)
        acc["est_global"].append(to_global_pose(est_rel, c))
        acc["mid_global"].append(to_global_pose(opt_rel, c))
        acc["gt"].append(gt[s:s + SEQ_LEN])
        res_g, info_g = glob.solve(opt_rel, h, eps[wi, 1])
        acc["opt_global"].append(to_global_pose(res_g.reshape(-1, 15, 3), c))
        infos.append((info_l, info_g))
    out = {k: merge_batches(np.asarray(v)) for k, v in acc.items()}
    if final_smooth:
        out["opt_global"] = gaussian_filter1d_reflect(out["opt_global"], 1.0, axis=0)
    out["infos"] = infos
    return out
