"""Drop-in mirror of the reference's optimiser surface (reference optimizer.py):

    BodyPoseOptimizer(camera_model_path, mean_skeleton, vae_path, seq_len, network_seq_len,
                      latent_dim, windows_size=5, overlap_size=1, slide_window=False, lr=2, max_iter=25)
        .set_weights(vae_weight, gmm_weight, smooth_weight, bone_length_weight, weight_3d, reproj_weight)
        .optimize_pose_seq_pytorch_LBFGS(relative_global_pose, heatmap_seq, smoothed_pose) -> (10,15,3) fp32
        .total_loss(hidden_parameter) -> scalar
    main(data_id, camera_model_path, vae_weight, gmm_weight, smoothness_weight, bone_length_weight,
         weight_3d, reproj_weight, visualization=False, final_smooth=False, merge=True, save=False,
         save_pose=False) -> (errors, final_estimated_seq, mid_local_pose_seq, final_optimized_seq, final_gt_seq)

Same names, argument meaning, pickle / checkpoint / camera formats and error behaviour
("norm is zero!").  What differs is the execution: every window of the clip is optimised in one
batch on the GPU through libgem_b200.so; there is no CPU path.  Arguments the reference accepts and
never reads (gmm_weight, merge, windows_size, slide_window, smoothed_pose, ...) are
accepted and ignored here too.  Extra keyword arguments (max_iter, eps, engine) are additions.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from .engine import Engine, GemError, energy_weights, heat_dims, heat_layout_code, lbfgs_params, tile_heat
from .metrics import calculate_errors
from .pipeline import SequenceOptimizer
from .vae_prep import PreparedVae

LOCAL_VAE_PATH = "networks/logs/only_local_full_dataset_latent_2048_len_10_kl_0.5_2/checkpoints/19.pth.tar"
GLOBAL_VAE_PATH = "networks/logs/real_full_dataset_latent_2048_len_10_slide_window_step_1_kl_0.5/checkpoints/19.pth.tar"
GEM_WIN_NORM_ZERO = 1
GEM_WIN_F16_RANGE = 2

_vae_cache = {}
_engine_cache = {}


def load_vae(vae_path, device, seq_len=10, latent_dim=2048):
    """torch.load(vae_path)['state_dict'] (optimizer.py:59), prepared once per (path, mtime):
    the reference re-reads both 130 MB checkpoints for every clip."""
    key = (os.path.abspath(vae_path), os.path.getmtime(vae_path), str(device), seq_len, latent_dim)
    if key not in _vae_cache:
        sd = torch.load(vae_path, map_location="cpu")["state_dict"]
        _vae_cache[key] = PreparedVae(sd, device, seq_len=seq_len, latent_dim=latent_dim)
    return _vae_cache[key]


def shared_engine(max_windows, max_history=24, latent_dim=2048, seq_len=10, heat_hw=(64, 64), num_joints=15):
    """One Engine per process (and heat-map geometry), grown when a clip needs more windows / a longer history."""
    key = (latent_dim, seq_len, tuple(heat_hw), num_joints,
           torch.cuda.current_device() if torch.cuda.is_available() else -1)
    eng = _engine_cache.get(key)
    if eng is None or eng.max_windows < max_windows or eng.max_history < max_history:
        if eng is not None:
            max_windows, max_history = max(max_windows, eng.max_windows), max(max_history, eng.max_history)
            eng.close()
        eng = Engine(max_windows=max(max_windows, 128), latent_dim=latent_dim, seq_len=seq_len,
                     num_joints=num_joints, heat_hw=tuple(heat_hw), max_history=max_history)
        _engine_cache[key] = eng
    return eng


def _raise_on_status(status):
    bits = 0
    if status.numel():
        for b in (GEM_WIN_NORM_ZERO, GEM_WIN_F16_RANGE):
            if int((status & b).sum()) != 0:
                bits |= b
    if bits & GEM_WIN_NORM_ZERO:
        raise Exception("norm is zero!")          # FishEyeCalibrated.py:124-127
    if bits & GEM_WIN_F16_RANGE:
        raise GemError("an activation or latent entry left fp16's finite range (65504): the split-fp16 tensor-core "
                       "operands saturated and the result is not fp32-faithful; re-run with "
                       "Engine.set_gemm_mode(1) (3xTF32, fp32 range)")


class BodyPoseOptimizer:
    kinematic_parents = [0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13]

    def __init__(self, camera_model_path, mean_skeleton, vae_path, seq_len, network_seq_len, latent_dim,
                 windows_size=5, overlap_size=1, slide_window=False, lr=2, max_iter=25, engine=None,
                 max_windows=128):
        self.seq_len, self.network_seq_len = seq_len, network_seq_len
        self.windows_size, self.slide_window, self.overlap_size = windows_size, slide_window, overlap_size
        self.lr, self.max_iter = lr, max_iter
        self.engine = engine if engine is not None else Engine(max_windows=max_windows, latent_dim=latent_dim,
                                                               seq_len=seq_len, max_history=max(max_iter - 1, 1))
        self.device = self.engine.device
        # mean over all frames of the bone lengths of `mean_skeleton` (optimizer.py:42-43, 89-94)
        sk = torch.as_tensor(np.asarray(mean_skeleton) if not isinstance(mean_skeleton, torch.Tensor)
                             else mean_skeleton).float().to(self.device).view(-1, 15, 3)
        bones = torch.linalg.vector_norm(sk - sk[:, self.kinematic_parents, :], dim=-1)
        self.mean_bone_length = torch.mean(bones, 0)
        self.engine.set_vae(0, load_vae(vae_path, self.device, seq_len, latent_dim))
        self.engine.set_camera_json(camera_model_path)
        self.vae_weight = self.gmm_weight = self.smooth_weight = None
        self.bone_length_weight = self.reproj_weight = self.weight_3d = None
        self.initial_pose = None
        self.heatmap_seq = None

    def set_weights(self, vae_weight, gmm_weight, smooth_weight, bone_length_weight, weight_3d, reproj_weight):
        self.vae_weight, self.gmm_weight, self.smooth_weight = vae_weight, gmm_weight, smooth_weight
        self.bone_length_weight, self.weight_3d, self.reproj_weight = bone_length_weight, weight_3d, reproj_weight

    def _weights(self):
        return energy_weights(self.weight_3d, self.smooth_weight, self.bone_length_weight, self.vae_weight,
                              self.reproj_weight)

    def optimize_windows(self, poses, heatmaps, eps=None, want_trace=False):
        """Batched form: poses [W,T,15,3], heatmaps [W,T,H,Wd,15] (HWC, as in the pickle)."""
        poses = torch.as_tensor(np.asarray(poses)).float()
        W = poses.shape[0]
        heat = fb = None
        if self.reproj_weight != 0:
            heat = torch.as_tensor(np.asarray(heatmaps)).float()
            heat = heat.reshape(W * self.seq_len, *heat.shape[-3:])
            fb = torch.arange(W, dtype=torch.int64) * self.seq_len
        if eps is None:
            eps = torch.randn(W, self.engine.n)               # one draw per stage call, SeqConvVAE.py:159-169
        self.engine.set_heat_layout(0)                        # [.., H, Wd, 15] as in the pickle (a shared engine may have been switched)
        res = self.engine.solve_stage(0, poses, heat, fb, torch.zeros(W, dtype=torch.int32),
                                      self.mean_bone_length, eps, self._weights(),
                                      lbfgs_params(lr=self.lr, max_iter=self.max_iter), want_trace=want_trace)
        _raise_on_status(res["status"])
        return res

    def optimize_pose_seq_pytorch_LBFGS(self, relative_global_pose, heatmap_seq, smoothed_pose, eps=None):
        res = self.optimize_windows(np.asarray(relative_global_pose)[None], np.asarray(heatmap_seq)[None],
                                    None if eps is None else np.asarray(eps).reshape(1, -1))
        self.initial_pose = torch.as_tensor(np.asarray(relative_global_pose)).float()
        if self.reproj_weight != 0:
            self.heatmap_seq = torch.as_tensor(np.asarray(heatmap_seq)).float()
        return res["pose"][0].cpu().numpy()

    # ---- the reference's per-term methods (optimizer.py:89-94, 139-149, 172-177, 202-218): same names and arguments,
    # each evaluated by the fused CUDA energy kernel (its un-weighted term sums) at the state the last optimize call left
    def calculate_bone_length(self, skeleton):
        """(T, 15*3) or (T, 15, 3) -> (T, 15) bone lengths; entry 0 is 0, the root being its own parent (optimizer.py:89-94)."""
        sk = torch.as_tensor(skeleton).to(self.device).view(-1, 15, 3)
        return torch.linalg.vector_norm(sk - sk[:, self.kinematic_parents, :], dim=-1)

    def _terms(self, x, need_anchor=False, need_heat=False):
        pose = torch.as_tensor(x).detach().float().reshape(1, self.seq_len, 15, 3)
        if need_anchor and self.initial_pose is None:
            raise GemError("pose_energy_3d needs the initial pose of a previous optimize_pose_seq_pytorch_LBFGS call")
        if need_heat and self.heatmap_seq is None:
            raise GemError("reprojection_energy_heatmap_fast needs the heat maps of a previous "
                           "optimize_pose_seq_pytorch_LBFGS call (with reproj_weight != 0)")
        anchor = self.initial_pose[None] if self.initial_pose is not None else pose
        heat = fb = None
        if need_heat:
            heat, fb = self.heatmap_seq, torch.zeros(1, dtype=torch.int64)
            self.engine.set_heat_layout(0)                      # this class holds the pickle's HWC maps
        _, terms, _, status = self.engine.energy_grad(pose, anchor, heat, fb, torch.zeros(1, dtype=torch.int32),
                                                      self.mean_bone_length,
                                                      energy_weights(1.0, 1.0, 1.0, 1.0, 1.0 if need_heat else 0.0))
        _raise_on_status(status)
        return terms[0]                      # {E_3d, E_smooth, E_bone, E_vae, E_reproj}, un-weighted

    def pose_energy_3d(self, x):
        """sum (x - initial_pose)^2 (optimizer.py:210-213)."""
        return self._terms(x, need_anchor=True)[0]

    def smooth_accelerate(self, x):
        """sum of squared second differences over the window's frames (optimizer.py:202-208)."""
        return self._terms(x)[1]

    def bone_length_energy(self, x):
        """sum (|x_j - x_parent(j)| - mean_bone_length_j)^2 (optimizer.py:172-177)."""
        return self._terms(x)[2]

    def vae_energy(self, hidden_parameter):
        """sum of squares of its argument — `total_loss` passes the DECODED POSE, not z (optimizer.py:215-218, 238)."""
        v = torch.as_tensor(hidden_parameter).detach().float()
        if v.numel() == self.seq_len * 45:
            return self._terms(v)[3]
        return torch.sum(torch.square(v.to(self.device)))       # any other shape: the one-liner itself

    def reprojection_energy_heatmap_fast(self, pose):
        """-sum of the heat maps sampled bilinearly at the fisheye projections of the joints (optimizer.py:139-149)."""
        return self._terms(pose, need_heat=True)[4]

    def total_loss(self, hidden_parameter):
        """E(z) at the state left by the last optimize call (initial pose / heatmaps), optimizer.py:226-240."""
        if self.initial_pose is None:
            raise GemError("total_loss needs a previous optimize_pose_seq_pytorch_LBFGS call (initial_pose unset)")
        z = torch.as_tensor(hidden_parameter).detach().float().reshape(1, -1)
        pose = self.engine.decode(0, z)
        heat = fb = None
        if self.reproj_weight != 0:
            heat, fb = self.heatmap_seq, torch.zeros(1, dtype=torch.int64)
            self.engine.set_heat_layout(0)
        E, _, _, status = self.engine.energy_grad(pose, self.initial_pose[None], heat, fb,
                                                  torch.zeros(1, dtype=torch.int32), self.mean_bone_length,
                                                  self._weights(), want_terms=False)
        _raise_on_status(status)
        return E[0]


CLIP_KEYS = ("estimated_local_skeleton", "gt_global_skeleton", "camera_pose_list", "heatmap_list")
_CLIP_DTYPES = {"estimated_local_skeleton": np.float64, "gt_global_skeleton": np.float64, "camera_pose_list": np.float64,
                "heatmap_list": np.float32}
_pinned_pool = {}


def _pinned(name, shape, np_dtype):
    """A pinned host array of the process-wide staging pool (grown when needed, reused across calls: page-locking
    gigabytes costs about as long as copying them).  Falls back to pageable memory without CUDA."""
    nbytes = int(np.prod(shape)) * np.dtype(np_dtype).itemsize
    if not torch.cuda.is_available():
        return torch.from_numpy(np.empty(shape, np_dtype))
    buf = _pinned_pool.get(name)
    if buf is None or buf.numel() < nbytes:
        _pinned_pool.pop(name, None)
        buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, pin_memory=True)
        _pinned_pool[name] = buf
    return buf[:nbytes].view({np.float32: torch.float32, np.float64: torch.float64}[np_dtype]).view(*shape)


class ClipSet(list):
    """The clips of one call: a list of per-clip dicts of (pinned) host tensors, plus `heat_all`, the ONE pinned tensor
    [total frames, H, W, J] (planar: [total frames, J, H, W]) every clip's heat maps are a view of — what the zero-copy
    path reads in place."""
    heat_all = None
    planar = 0          # layout code of the heat maps: 0 the pickle's HWC, 1 planar, 2 tiled (engine.heat_layout_code)


def _unpickle(data_id):
    with open("{}/test_data.pkl".format(data_id), "rb") as f:
        return pickle.load(f)


def load_clips(data_ids, pinned=True, planar="tiled"):
    """'<data_id>/test_data.pkl' of every clip (optimizer.py:315-324; extra keys are ignored) unpickled and landed
    ONCE in page-locked host memory: each per-frame list is stacked straight into its slice of one pinned buffer
    per key, so the solver can read the heat maps where they are (or DMA them) without another host copy.

    planar=True: the heat maps are stacked as [frames, J, H, W] — the permutation the reference applies to every window
    before grid_sample (optimizer.py:251), done here once while the frames are copied anyway.  With planar maps the
    x-neighbours of a bilinear footprint share a sector, which is what makes the zero-copy path cheap;
    planar="tiled" (or 2; the default, for maps with H % 4 == 0 and W % 8 == 0, else plain planar) goes one step
    further and stores every map as [H/4, W/8] tiles of 4 rows x 8 texels, one 128-byte line each, so that ONE PCIe
    request brings a joint's two-dimensional neighbourhood (the bus charges per request: DESIGN.md section 4);
    planar=False keeps the pickle's [frames, H, W, J]."""
    layout = heat_layout_code(planar)
    raws = [_unpickle(d) for d in data_ids]
    n_frames = [len(r["estimated_local_skeleton"]) for r in raws]
    offs = np.concatenate([[0], np.cumsum(n_frames)]).astype(np.int64)
    total = int(offs[-1])
    clips = ClipSet({} for _ in raws)
    for key in CLIP_KEYS:
        if any(key not in r for r in raws):
            if key == "gt_global_skeleton":
                continue
            raise KeyError("test_data.pkl lacks the key {!r}".format(key))
        shape = tuple(np.asarray(next(r[key][0] for r in raws if len(r[key]))).shape) if total else ()
        to_planar = layout if key == "heatmap_list" and len(shape) == 3 else 0
        if to_planar == 2 and (shape[0] % 4 or shape[1] % 8):
            to_planar = 1                              # maps that do not tile: plain planar
        if to_planar == 2:
            shape = (shape[2], shape[0] // 4, shape[1] // 8, 4, 8)
        elif to_planar:
            shape = (shape[2], shape[0], shape[1])
        dt = _CLIP_DTYPES[key]
        buf = _pinned("clip_" + key, (total,) + shape, dt) if pinned else torch.from_numpy(np.empty((total,) + shape, dt))
        view = buf.numpy()
        for i, r in enumerate(raws):
            frames = r[key]
            if len(frames) != n_frames[i]:
                raise ValueError("{}: {} has {} frames, expected {}".format(data_ids[i], key, len(frames), n_frames[i]))
            dst = view[offs[i]:offs[i + 1]]
            if n_frames[i] and to_planar:
                _stack_planar(frames, dst, dt, to_planar)
            elif n_frames[i]:
                if isinstance(frames, np.ndarray):
                    np.copyto(dst, frames, casting="same_kind")
                else:
                    np.stack(frames, out=dst, casting="same_kind")
            clips[i][key] = buf[offs[i]:offs[i + 1]]
        if key == "heatmap_list":
            clips.heat_all = buf
            clips.planar = int(to_planar)
    return clips


_STACK_CHUNK = 64          # frames per piece: 15.7 MB of 64 x 64 x 15 maps, small enough to stay in the host's caches


def _stack_planar(frames, dst, dt, layout):
    """HWC per-frame maps -> their planar (layout 1) or tiled (2) slice `dst` of the staging buffer, piece by piece: a
    piece is stacked into a scratch array that stays in cache and leaves through ONE strided copy, and the pieces run
    on a few host threads (numpy and torch release the GIL while they copy).  Stacking the whole list first costs a
    fresh allocation the size of the clip and a second pass through DRAM (0.45 s against 0.10 s for 1500 frames on
    eight cores)."""
    from concurrent.futures import ThreadPoolExecutor
    n = len(frames)
    out = torch.from_numpy(dst)

    def piece(lo):
        hi = min(lo + _STACK_CHUNK, n)
        src = frames[lo:hi] if isinstance(frames, np.ndarray) else np.stack(frames[lo:hi])
        src = torch.from_numpy(np.ascontiguousarray(src, dtype=dt)).permute(0, 3, 1, 2)
        out[lo:hi].copy_(tile_heat(src) if layout == 2 else src)

    starts = range(0, n, _STACK_CHUNK)
    workers = min(8, os.cpu_count() or 1, len(starts))
    if workers <= 1:
        for lo in starts:
            piece(lo)
    else:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(piece, starts))


def load_clip(data_id, pinned=True):
    """'<data_id>/test_data.pkl' -> dict of (pinned) host tensors (optimizer.py:315-324); extra keys are ignored."""
    return load_clips([data_id], pinned=pinned)[0]


def _shape_of(frames):
    """Shape of a per-frame array / tensor / list of per-frame arrays without stacking the list."""
    if hasattr(frames, "shape"):
        return tuple(frames.shape)
    return (len(frames),) + (tuple(np.shape(frames[0])) if len(frames) else ())


def _n_windows(n_frames, seq_len=10, overlap=2):
    return len(range(0, n_frames - seq_len + 1, seq_len - overlap))


def solve_clips(clips, camera_model_path, vae_weight=0.0, gmm_weight=0.0, smoothness_weight=0.001, bone_length_weight=0.01,
                weight_3d=0.01, reproj_weight=0.01, final_smooth=False, max_iter=25, eps=None,
                local_vae_path=LOCAL_VAE_PATH, global_vae_path=GLOBAL_VAE_PATH, engine=None, ingest="auto",
                outputs="all", copy_stream=None, group=None):
    """The batched path behind `main` / `main_batch`, for clips that are already in memory: every window of every clip
    through local stage -> SLAM transform -> global stage -> stitching in ONE batched solve.  Nothing here
    synchronises with the host; the results are device tensors.

    clips   : list of dicts (reference pickle keys) of numpy arrays / host tensors / CUDA tensors, a `ClipSet` from
              `load_clips`, or an already built `pipeline.WindowBatch` (inputs resident in HBM).
    ingest  : how host heat maps reach the GPU.  "zero_copy": they stay in ONE pinned host tensor (`ClipSet.heat_all`)
              and the energy kernel fetches the texels it samples over PCIe (about 2 % of the maps);
              "upload": piecewise host-to-device copies on `copy_stream`, the solve starts on the first pieces while the
              rest is in flight; "auto": zero_copy when `clips.heat_all` is pinned, else upload.
    outputs : "all" = the six stitched sequences of optimizer.py:442-450 per clip; "optimized" = only the final
              optimised global sequence (with torch.distributed initialised: this rank's frames of it, the clips'
              windows being sharded over the ranks; `globalegomocap_b200.distributed.stitch_optimized`).
    Returns dict(batch, sol, merged); `merged[i]` is a dict of device tensors for clip i (None for a clip shorter
    than one window)."""
    from . import distributed as gdist
    from .pipeline import WindowBatch
    if ingest not in ("auto", "zero_copy", "upload"):
        raise ValueError("ingest must be 'auto', 'zero_copy' or 'upload'")
    if outputs not in ("all", "optimized"):
        raise ValueError("outputs must be 'all' or 'optimized'")
    prebuilt = isinstance(clips, WindowBatch)
    planar = heat_layout_code(getattr(clips, "planar", 0))
    if prebuilt:
        batch_W = clips.W
        heat_shape = heat_dims(clips.heat.shape, planar)
    else:
        batch_W = sum(_n_windows(len(c["estimated_local_skeleton"])) for c in clips)
        heat_shape = heat_dims(_shape_of(clips[0]["heatmap_list"]), planar) if len(clips) else (64, 64, 15)
    eng = engine if engine is not None else shared_engine(batch_W, max(max_iter - 1, 1), heat_hw=heat_shape[:2],
                                                          num_joints=heat_shape[2])
    eng.set_camera_json(camera_model_path) if isinstance(camera_model_path, str) else eng.set_camera(*camera_model_path)
    for which, path in ((0, local_vae_path), (1, global_vae_path)):
        eng.set_vae(which, load_vae(path, eng.device) if isinstance(path, str) else path)
    seq_opt = SequenceOptimizer(eng, vae_weight=vae_weight, smoothness_weight=smoothness_weight,
                                bone_length_weight=bone_length_weight, weight_3d=weight_3d,
                                reproj_weight=reproj_weight, lr=2, max_iter=max_iter)
    if prebuilt:
        batch = clips
    else:
        heat_all = getattr(clips, "heat_all", None)
        zero_copy_ok = (heat_all is not None and isinstance(heat_all, torch.Tensor) and not heat_all.is_cuda and
                        heat_all.is_pinned() and reproj_weight != 0)
        if ingest == "zero_copy" and not zero_copy_ok:
            raise ValueError("ingest='zero_copy' needs the clips' heat maps in one pinned host tensor (load_clips)")
        if zero_copy_ok and ingest in ("auto", "zero_copy"):
            batch = WindowBatch(eng, clips, host_heat=heat_all, planar=planar)
        else:
            host_side = len(clips) > 0 and not (isinstance(clips[0]["heatmap_list"], torch.Tensor) and
                                                clips[0]["heatmap_list"].is_cuda)
            if host_side and copy_stream is None:
                copy_stream = _copy_stream(eng.device)
            batch = WindowBatch(eng, clips, copy_stream=copy_stream if host_side else None, planar=planar)
    sol = seq_opt.solve(batch, eps=eps)
    if batch.ready_events:
        eng.set_slices(None)
    if outputs == "all":
        merged = seq_opt.stitch(batch, sol, final_smooth=final_smooth is True)
    else:
        opt_global = eng.to_global(sol["glob"]["pose"], sol["cams"])                  # optimizer.py:421
        wo = batch.window_offsets
        live = [i for i, n in enumerate(batch.n_windows) if n > 0]
        seqs = gdist.stitch_optimized(eng, [opt_global[wo[i]:wo[i + 1]] for i in live],
                                      final_smooth=final_smooth is True, group=group)
        merged = [None] * len(batch.n_windows)
        for i, q in zip(live, seqs):
            merged[i] = {"final_optimized_seq": q}
    return dict(batch=batch, sol=sol, merged=merged, engine=eng)


_copy_streams = {}


def _copy_stream(device):
    key = str(device)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def main_batch(data_ids, camera_model_path, vae_weight, gmm_weight, smoothness_weight, bone_length_weight, weight_3d,
               reproj_weight, visualization=False, final_smooth=False, merge=True, save=False, save_pose=False,
               max_iter=25, eps=None, local_vae_path=LOCAL_VAE_PATH, global_vae_path=GLOBAL_VAE_PATH, engine=None,
               ingest="auto"):
    """`main` for several clips at once: every window of every clip is optimised in ONE batched solve
    instead of one `main` call per 100-frame clip (optimize_whole_sequence.py:48-63 loops over the clips, and
    each call reloads both checkpoints, optimizer.py:332-350).  Returns the list of `main`'s return tuples, clip by
    clip.  The pickles are landed once in pinned host memory (`load_clips`) and the heat maps are read from there by
    the GPU (`solve_clips`, ingest).  The reparameterisation noise is drawn as one (sum W, 2, latent) block, which
    consumes the global torch generator exactly like consecutive `main` calls do, so batched and per-clip runs see the
    same z0."""
    if visualization is True or save:
        # (checked before any work is done: the reference's open3d mesh export is outside this path)
        raise NotImplementedError("mesh visualisation / .ply export need open3d and are outside this path")
    clips = load_clips(data_ids)
    n_win = [_n_windows(len(c["estimated_local_skeleton"])) for c in clips]
    for d, n in zip(data_ids, n_win):
        if n <= 0:
            raise ValueError("clip {} is shorter than one window ({} frames)".format(d, 10))
    if "gt_global_skeleton" not in clips[0]:
        raise KeyError("test_data.pkl lacks the key 'gt_global_skeleton' (optimizer.py:318)")
    if eps is None:
        # the reference draws torch.randn_like(std) of shape (1, 2048) in the order local(w0),
        # global(w0), local(w1), ... from the global generator; one CPU draw of (W, 2, 2048)
        # consumes the same stream
        eps = torch.randn(sum(n_win), 2, 2048)
    out = solve_clips(clips, camera_model_path, vae_weight, gmm_weight, smoothness_weight, bone_length_weight, weight_3d,
                      reproj_weight, final_smooth=final_smooth, max_iter=max_iter, eps=eps,
                      local_vae_path=local_vae_path, global_vae_path=global_vae_path, engine=engine, ingest=ingest)
    _raise_on_status(out["sol"]["local"]["status"])
    results = []
    for data_id, m in zip(data_ids, out["merged"]):
        final_estimated_seq = list(m["final_estimated_seq"].cpu().numpy())
        mid_estimated_seq = list(m["mid_estimated_seq"].cpu().numpy())
        mid_local_pose_seq = list(m["mid_local_pose_seq"].cpu().numpy())
        final_gt_seq = list(m["final_gt_seq"].cpu().numpy())
        final_optimized_seq = m["final_optimized_seq"].cpu().numpy()
        if final_smooth is not True:
            final_optimized_seq = list(final_optimized_seq)
        if save_pose:
            dataset_dir, seq_name = os.path.split(data_id)
            dataset_name = os.path.split(dataset_dir)[1]
            out_dir = "out/{}/{}".format(dataset_name, seq_name)
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, "result_pose.pkl"), "wb") as f:
                pickle.dump({"estimated_pose": final_estimated_seq, "optimized_pose": final_optimized_seq,
                             "mid_optimized_pose": mid_estimated_seq, "gt_pose": final_gt_seq}, f)
        errors = calculate_errors(final_estimated_seq, mid_estimated_seq, final_optimized_seq, final_gt_seq,
                                  on_device=True)
        results.append((errors, final_estimated_seq, mid_local_pose_seq, final_optimized_seq, final_gt_seq))
    return results


def main(data_id, camera_model_path, vae_weight, gmm_weight, smoothness_weight, bone_length_weight, weight_3d,
         reproj_weight, visualization=False, final_smooth=False, merge=True, save=False, save_pose=False,
         max_iter=25, eps=None, local_vae_path=LOCAL_VAE_PATH, global_vae_path=GLOBAL_VAE_PATH, engine=None,
         ingest="auto"):
    """The reference's `optimizer.main` (optimizer.py:311-507) for one clip."""
    return main_batch([data_id], camera_model_path, vae_weight, gmm_weight, smoothness_weight, bone_length_weight,
                      weight_3d, reproj_weight, visualization=visualization, final_smooth=final_smooth, merge=merge,
                      save=save, save_pose=save_pose, max_iter=max_iter, eps=eps, local_vae_path=local_vae_path,
                      global_vae_path=global_vae_path, engine=engine, ingest=ingest)[0]
