"""Deterministic synthetic inputs for the pose-sequence optimiser.

Everything here is generated with numpy's PCG64 stream and IEEE-exact
arithmetic (+, -, *, /, sqrt) wherever possible, so that the CPU-only container
that produced ``tests/golden/*.npz`` and the GPU box that replays them build
the same inputs without shipping multi-megabyte fixtures.

Shapes follow the reference's per-sequence pickle (``test_data.pkl``, read at
reference ``optimizer.py:315-324``, written by
``MakeDataForOptimization/process_test_data.py:149-157``):

    estimated_local_skeleton : list of N arrays (15, 3) float64, camera frame, metres
    gt_global_skeleton       : list of N arrays (15, 3) float64
    camera_pose_list         : list of N arrays (4, 4)  float64, camera-to-world (SLAM)
    heatmap_list             : list of N arrays (64, 64, 15) float32, HWC

and the VAE checkpoint layout of reference ``networks/models/SeqConvVAE.py``
(``torch.load(path)['state_dict']``, reference ``optimizer.py:59-60``).
"""
from __future__ import annotations

import json
import os
import pickle
from collections import OrderedDict

import numpy as np

KINEMATIC_PARENTS = (0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13)  # optimizer.py:34
NUM_JOINTS = 15
SEQ_LEN = 10
OVERLAP = 2
LATENT_DIM = 2048
HEATMAP_SIZE = 64

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_CAMERA_JSON = os.path.join(_DATA_DIR, "fisheye.calibration.json")


# --------------------------------------------------------------------------
# camera (float64 numpy; only used to place the synthetic heatmap peaks)
# --------------------------------------------------------------------------
def load_camera(path: str = DEFAULT_CAMERA_JSON):
    """Returns (poly_w2c float64[P], cx, cy) from a calibration JSON with the
    reference's keys (reference FishEyeCalibrated.py:8-14)."""
    with open(path) as f:
        cal = json.load(f)
    intr = np.asarray(cal["intrinsic"], dtype=np.float64)
    poly = np.asarray(cal["polynomialW2C"], dtype=np.float64)
    return poly, float(intr[0][2]), float(intr[1][2])


def project_fisheye_np(points: np.ndarray, poly: np.ndarray, cx: float, cy: float) -> np.ndarray:
    """float64 omnidirectional projection, (…,3) -> (…,2). Same model as the
    reference's world2camera (FishEyeCalibrated.py:59-88)."""
    p = np.asarray(points, dtype=np.float64)
    x, y, z = p[..., 0], p[..., 1], -p[..., 2]
    r = np.sqrt(x * x + y * y)
    theta = np.arctan(z / r)
    rho = np.full_like(theta, poly[0])
    t_i = np.ones_like(theta)
    for c in poly[1:]:
        t_i = t_i * theta
        rho = rho + t_i * c
    return np.stack([x / r * rho + cx, y / r * rho + cy], axis=-1)


# --------------------------------------------------------------------------
# VAE weights
# --------------------------------------------------------------------------
_ENC_DIMS = (64, 64, 128, 256, 512)
_DEC_DIMS = (512, 256, 128, 64, 64)


def vae_state_dict_shapes(latent_dim: int = LATENT_DIM, seq_len: int = SEQ_LEN, channels: int = 45):
    """Key -> shape of ConvVAE(in=45, out=45, latent_dim, seq_len).state_dict()
    (reference SeqConvVAE.py:27-92; order matches nn.Module registration)."""
    shapes = OrderedDict()
    cin = channels
    for i, h in enumerate(_ENC_DIMS):
        shapes[f"encoder.{i}.0.weight"] = (h, cin, 3)
        shapes[f"encoder.{i}.0.bias"] = (h,)
        for k in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"encoder.{i}.1.{k}"] = (h,)
        shapes[f"encoder.{i}.1.num_batches_tracked"] = ()
        cin = h
    flat = _ENC_DIMS[-1] * seq_len
    shapes["fc_mu.weight"] = (latent_dim, flat)
    shapes["fc_mu.bias"] = (latent_dim,)
    shapes["fc_var.weight"] = (latent_dim, flat)
    shapes["fc_var.bias"] = (latent_dim,)
    shapes["decoder_input.weight"] = (flat, latent_dim)
    shapes["decoder_input.bias"] = (flat,)
    for i in range(len(_DEC_DIMS) - 1):
        shapes[f"decoder.{i}.0.weight"] = (_DEC_DIMS[i], _DEC_DIMS[i + 1], 3)  # ConvT: (in,out,k)
        shapes[f"decoder.{i}.0.bias"] = (_DEC_DIMS[i + 1],)
        for k in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"decoder.{i}.1.{k}"] = (_DEC_DIMS[i + 1],)
        shapes[f"decoder.{i}.1.num_batches_tracked"] = ()
    c = _DEC_DIMS[-1]
    shapes["final_layer.0.weight"] = (c, c, 3)
    shapes["final_layer.0.bias"] = (c,)
    for k in ("weight", "bias", "running_mean", "running_var"):
        shapes[f"final_layer.1.{k}"] = (c,)
    shapes["final_layer.1.num_batches_tracked"] = ()
    shapes["final_layer.3.weight"] = (channels, c, 3)
    shapes["final_layer.3.bias"] = (channels,)
    return shapes


def make_vae_state_dict(seed: int, *, perturb_bn: bool = True, pose_bias: np.ndarray | None = None,
                        latent_dim: int = LATENT_DIM, seq_len: int = SEQ_LEN, gain: float = 1.0):
    """Random-init checkpoint as a dict of numpy arrays (fp32; int64 counters).

    Weights/biases ~ U(-b, b), b = gain/sqrt(fan_in) (PyTorch's default bound).
    ``perturb_bn`` draws non-trivial BatchNorm affine + running statistics so BN
    folding is exercised; ``pose_bias`` (45,) replaces the last conv's bias so
    decoded poses sit near a plausible skeleton (joints then land on the
    heatmaps and the reprojection term is live).
    """
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for key, shape in vae_state_dict_shapes(latent_dim, seq_len).items():
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[key] = np.asarray(0, dtype=np.int64)
            continue
        is_bn = ".1." in key and len(shape) == 1 and not key.startswith("fc_")
        if is_bn:
            if not perturb_bn:
                val = {"weight": 1.0, "bias": 0.0, "running_mean": 0.0, "running_var": 1.0}[leaf]
                sd[key] = np.full(shape, val, dtype=np.float32)
            else:
                lo, hi = {"weight": (0.5, 1.5), "bias": (-0.1, 0.1),
                          "running_mean": (-0.1, 0.1), "running_var": (0.5, 1.5)}[leaf]
                sd[key] = rng.uniform(lo, hi, size=shape).astype(np.float32)
            continue
        wkey = key.rsplit(".", 1)[0] + ".weight"
        wshape = vae_state_dict_shapes(latent_dim, seq_len)[wkey]
        fan_in = wshape[1] * (wshape[2] if len(wshape) == 3 else 1)
        b = gain / np.sqrt(float(fan_in))
        sd[key] = rng.uniform(-b, b, size=shape).astype(np.float32)
    if pose_bias is not None:
        sd["final_layer.3.bias"] = np.asarray(pose_bias, dtype=np.float32).reshape(45).copy()
    return sd


def save_checkpoint(state_dict, path: str):
    """Writes the reference's checkpoint format: torch.save({'state_dict': ...})
    (reference networks/train.py:102-108, read at optimizer.py:59)."""
    import torch

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"state_dict": OrderedDict((k, torch.from_numpy(np.array(v))) for k, v in state_dict.items())},
               path)


# --------------------------------------------------------------------------
# clips
# --------------------------------------------------------------------------
def _wave(phase: np.ndarray) -> np.ndarray:
    """Smooth period-1 wave in [-1, 1] from exact ops (parabolic sine)."""
    f = phase - np.floor(phase)
    half = f < 0.5
    g = np.where(half, f, f - 0.5)
    lobe = 16.0 * g * (0.5 - g)
    return np.where(half, lobe, -lobe)


def make_clip(n_frames: int, seed: int, *, camera_json: str = DEFAULT_CAMERA_JSON,
              heat_radius: float = 5.0, est_noise: float = 0.01, hm: int = HEATMAP_SIZE,
              as_lists: bool = False):
    """One synthetic sequence in the reference's pickle layout.

    Base pose ~ U([-0.4,-0.4,0.2],[0.4,0.4,1.3]) per joint (camera frame, z > 0
    is in front of the fisheye), smooth per-coordinate motion of amplitude 5 cm,
    heatmap peak at the reference's pixel mapping ((u-128)*63/1024, v*63/1024)
    of the true pose (optimizer.py:143-147), compact quartic bump of radius
    ``heat_radius`` px, cameras with slow yaw + drift.
    """
    rng = np.random.default_rng(seed)
    poly, cx, cy = load_camera(camera_json)
    base = rng.uniform([-0.4, -0.4, 0.2], [0.4, 0.4, 1.3], size=(NUM_JOINTS, 3))
    phase = rng.uniform(0.0, 1.0, size=(NUM_JOINTS, 3))
    tau = (np.arange(n_frames, dtype=np.float64) / max(n_frames, 1))[:, None, None]
    true_local = base[None] + 0.05 * _wave(3.0 * tau * (n_frames / 100.0) + phase[None])
    est_local = true_local + est_noise * rng.standard_normal(size=true_local.shape)

    # cameras: rational rotation about the z axis (exact ops), slow drift
    i = np.arange(n_frames, dtype=np.float64)
    u = 0.15 * i / max(n_frames, 1)
    c, s = (1 - u * u) / (1 + u * u), 2 * u / (1 + u * u)
    cams = np.zeros((n_frames, 4, 4), dtype=np.float64)
    cams[:, 0, 0], cams[:, 0, 1] = c, -s
    cams[:, 1, 0], cams[:, 1, 1] = s, c
    cams[:, 2, 2] = 1.0
    cams[:, 3, 3] = 1.0
    cams[:, 0, 3] = 0.5 * i / max(n_frames, 1)
    cams[:, 1, 3] = 0.1 * _wave(i / 63.0)
    cams[:, 2, 3] = 0.02 * i / max(n_frames, 1)

    gt_local = true_local + 0.02 * rng.standard_normal(size=true_local.shape)
    gt_global = np.einsum("nij,nkj->nki", cams[:, :3, :3], gt_local) + cams[:, None, :3, 3]

    uv = project_fisheye_np(true_local, poly, cx, cy)                # (N,15,2)
    px = (uv[..., 0] - 128.0) * (hm - 1) / 1024.0
    py = uv[..., 1] * (hm - 1) / 1024.0
    grid = np.arange(hm, dtype=np.float64)
    inv_r2 = 1.0 / (heat_radius * heat_radius)
    bx = np.maximum(0.0, 1.0 - (grid[None, None, :] - px[..., None]) ** 2 * inv_r2)  # (N,15,W)
    by = np.maximum(0.0, 1.0 - (grid[None, None, :] - py[..., None]) ** 2 * inv_r2)  # (N,15,H)
    heat = np.einsum("njy,njx->nyxj", by * by, bx * bx).astype(np.float32)           # (N,H,W,15)

    clip = {
        "estimated_local_skeleton": est_local,
        "gt_global_skeleton": gt_global,
        "camera_pose_list": cams,
        "heatmap_list": heat,
    }
    if as_lists:
        clip = {k: [a for a in v] for k, v in clip.items()}
    return clip


def dense_heat_window(seed: int = 0, frames: int = SEQ_LEN, hm: int = HEATMAP_SIZE) -> np.ndarray:
    """(frames, hm, hm, 15) fp32 dense pattern from integer arithmetic only
    (non-zero right up to the borders; used for the zero-padding edge cases)."""
    t = np.arange(frames, dtype=np.int64)[:, None, None, None]
    y = np.arange(hm, dtype=np.int64)[None, :, None, None]
    x = np.arange(hm, dtype=np.int64)[None, None, :, None]
    j = np.arange(NUM_JOINTS, dtype=np.int64)[None, None, None, :]
    v = (131 * x + 373 * y + 977 * j + 613 * t + 7919 * (seed + 1) + 17 * x * y + 29 * j * t) % 1021
    return (v.astype(np.float64) / 1021.0).astype(np.float32)


def point_for_heat_pixel(ix: float, iy: float, r: float, poly, cx, cy, hm: int = HEATMAP_SIZE) -> np.ndarray:
    """3-D camera-frame point at horizontal distance ``r`` from the optical axis whose fast-path heatmap coordinate
    (optimizer.py:143-147: ix = (u-128)*(hm-1)/1024, iy = v*(hm-1)/1024) is (ix, iy)."""
    u = ix * 1024.0 / (hm - 1) + 128.0
    v = iy * 1024.0 / (hm - 1)
    du, dv = u - cx, v - cy
    rho_t = float(np.hypot(du, dv))

    def rho(th):
        acc, ti = poly[0], 1.0
        for c in poly[1:]:
            ti *= th
            acc += ti * c
        return acc

    lo, hi = -np.pi / 2 + 1e-6, 0.6         # rho is increasing in theta on this range
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if rho(mid) < rho_t:
            lo = mid
        else:
            hi = mid
    th = 0.5 * (lo + hi)
    return np.array([r * du / rho_t, r * dv / rho_t, -r * np.tan(th)], dtype=np.float64)


def write_clip_pickle(clip, directory: str):
    """Writes ``<directory>/test_data.pkl`` as the reference's data-prep does
    (process_test_data.py:149-157): a dict of python lists of per-frame arrays."""
    os.makedirs(directory, exist_ok=True)
    out = {k: [np.asarray(a) for a in v] for k, v in clip.items()}
    with open(os.path.join(directory, "test_data.pkl"), "wb") as f:
        pickle.dump(out, f)
    return os.path.join(directory, "test_data.pkl")


def mean_pose_bias(clip) -> np.ndarray:
    """(45,) bias placing a random decoder's output near the clip's mean pose."""
    return np.asarray(clip["estimated_local_skeleton"], dtype=np.float64).mean(axis=0).reshape(45).astype(np.float32)


def mean_bone_length(skeleton) -> np.ndarray:
    """(15,) fp32 mean bone lengths over all frames of a clip's local estimate, cast to fp32 first
    (BodyPoseOptimizer.__init__, reference optimizer.py:42-43, 89-94, 333)."""
    s = np.asarray(skeleton).astype(np.float32).reshape(-1, NUM_JOINTS, 3)
    b = s - s[:, list(KINEMATIC_PARENTS), :]
    return np.sqrt((b * b).sum(-1)).mean(0, dtype=np.float32)


def window_starts(n_frames: int, seq_len: int = SEQ_LEN, overlap: int = OVERLAP):
    """Bit-exact window partition of reference optimizer.py:370."""
    return list(range(0, n_frames - seq_len + 1, seq_len - overlap))
