"""Batched per-sequence pipeline: every window of every clip goes through the
local stage, the SLAM camera transform and the global stage together, then the
clips are stitched (reference optimizer.py:370-450, which loops window by window).

Host logic only: window partition, index bookkeeping and buffer placement.  All
arithmetic runs in libgem_b200.so through ``Engine``.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import KINEMATIC_PARENTS, Engine, energy_weights, heat_layout_code, lbfgs_params

SEQ_LEN = 10
OVERLAP = 2


def window_starts(n_frames, seq_len=SEQ_LEN, overlap=OVERLAP):
    """range(0, N - seq_len + 1, seq_len - overlap), reference optimizer.py:370: the trailing
    (N - seq_len) mod (seq_len - overlap) frames are dropped."""
    return list(range(0, n_frames - seq_len + 1, seq_len - overlap))


def stage_weights(vae_weight, smoothness_weight, bone_length_weight, weight_3d, reproj_weight):
    """(local, global) energy weights exactly as main wires them (optimizer.py:352-358): the global
    stage hard-codes bone_length_weight=0.01 and reproj_weight=0, the local stage scales the
    smoothness weight by 1/100 and weight_3d by 1/10000."""
    local = energy_weights(weight_3d / 10000, smoothness_weight / 100, bone_length_weight, vae_weight, reproj_weight)
    glob = energy_weights(weight_3d, smoothness_weight, 0.01, vae_weight, 0)
    return local, glob


def upload_pieces(n_windows, n_frames, seq_len=SEQ_LEN, overlap=OVERLAP, min_piece_windows=96, max_pieces=4):
    """Cuts one clip into upload pieces of whole windows: [(first window, end window, first frame, end frame)].
    A clip is cut while its pieces keep >= min_piece_windows windows (at most max_pieces); piece starts are even
    window indices (the energy kernel pairs windows).  Piece = windows [a, b) -> frames up to stride*(b-1)+seq_len;
    the next piece starts where this one ended, the last one takes the clip's trailing frames: the frame ranges
    are disjoint and cover the clip."""
    if n_windows <= 0:
        return []
    stride = seq_len - overlap
    n_pieces = max(1, min(max_pieces, n_windows // max(min_piece_windows, 1)))
    cuts = [(n_windows * k // n_pieces) // 2 * 2 for k in range(n_pieces)] + [n_windows]
    out, f_prev = [], 0
    for k in range(n_pieces):
        a, b = cuts[k], cuts[k + 1]
        f_end = n_frames if k == n_pieces - 1 else stride * (b - 1) + seq_len
        out.append((a, b, f_prev, f_end))
        f_prev = f_end
    return out


class WindowBatch:
    """Device-resident inputs of all windows of a list of clips.

    With ``copy_stream`` the host-to-device copies run on that stream, clip after clip, and every clip's
    heat maps get a completion event: ``SequenceOptimizer.solve`` hands the events to the library, which
    starts optimising the first clips while the later ones are still crossing PCIe (the clips become the
    library's slices).

    With ``host_heat`` (ONE pinned float32 tensor [total frames, H, W, J] holding the clips' heat maps back to
    back) the maps are not copied at all: the energy kernel reads them over PCIe through its texel cache, so
    only the texels the optimiser samples ever cross the bus."""

    def __init__(self, engine: Engine, clips, copy_stream=None, min_piece_windows=96, host_heat=None, planar=False):
        # planar: layout of the clips' heat maps (and of host_heat): False / 0 the pickle's [frames, H, W, J], True / 1
        # [frames, J, H, W], 2 / "tiled" [frames, J, H/4, W/8, 4, 8] (optimizer.load_clips); SequenceOptimizer.solve tells
        # the engine
        dev = engine.device
        self.planar = heat_layout_code(planar)     # 0 HWC, 1 planar, 2 tiled
        if len(clips):
            h0 = clips[0]["heatmap_list"] if host_heat is None else host_heat
            ndim = h0.ndim if hasattr(h0, "ndim") else (1 + np.ndim(h0[0]) if len(h0) else (6 if self.planar == 2 else 4))
            if ndim != (6 if self.planar == 2 else 4):
                raise ValueError("heat maps with {} dimensions do not fit the declared layout {} ([frames, H, W, J], "
                                 "[frames, J, H, W] or tiled [frames, J, H/4, W/8, 4, 8])".format(ndim, self.planar))
        self.n_frames = [len(c["estimated_local_skeleton"]) for c in clips]
        self.starts = [window_starts(n, engine.T, OVERLAP) for n in self.n_frames]
        self.n_windows = [len(s) for s in self.starts]
        offs = np.concatenate([[0], np.cumsum(self.n_frames)])
        self.frame_offsets = offs
        fb = np.concatenate([np.asarray(s, dtype=np.int64) + offs[i] for i, s in enumerate(self.starts)]
                            or [np.zeros(0, np.int64)])
        self.W = int(fb.shape[0])
        clip_idx = np.concatenate([np.full(len(s), i, dtype=np.int32) for i, s in enumerate(self.starts)]
                                  or [np.zeros(0, np.int32)])
        self.window_offsets = np.concatenate([[0], np.cumsum(self.n_windows)]).astype(np.int64)
        self.ready_events, self.ready_first_windows = [], []

        # The window index tensors depend only on the clips' lengths: built once per (lengths, device) and kept with the
        # engine.  A pageable host-to-device copy per call would also synchronise the host with everything already
        # enqueued, so that consecutive calls could not overlap their host-side preparation with the previous solve.
        cache = engine.__dict__.setdefault("_index_cache", {})
        key = (tuple(self.n_frames), engine.T, str(dev))
        if key not in cache:
            if len(cache) > 16:
                cache.clear()
            frame_base = torch.as_tensor(fb, device=dev)
            cache[key] = (frame_base, torch.as_tensor(clip_idx, device=dev), torch.as_tensor(KINEMATIC_PARENTS, device=dev),
                          frame_base[:, None] + torch.arange(engine.T, device=dev)[None, :])      # [W,T] frame ids
        self.frame_base, self.clip_idx, parents, self.frame_idx = cache[key]

        total = int(offs[-1])

        def src_of(c, key):
            return c[key] if isinstance(c[key], torch.Tensor) else torch.as_tensor(np.asarray(c[key]))

        def alloc(key, dtype):
            first = src_of(clips[0], key)
            return torch.empty((total,) + tuple(first.shape[1:]), dtype=dtype, device=dev)

        def fill(buf, key, clip_range):
            # each clip is copied straight into its slice (asynchronously when the source is a pinned tensor)
            for i in clip_range:
                buf[offs[i]:offs[i + 1]].copy_(src_of(clips[i], key), non_blocking=True)

        has_gt = "gt_global_skeleton" in clips[0]
        self.est = alloc("estimated_local_skeleton", torch.float64)        # [F,15,3]
        self.cams = alloc("camera_pose_list", torch.float64)               # [F,4,4]
        self.gt = alloc("gt_global_skeleton", torch.float64) if has_gt else None
        every = range(len(clips))
        if host_heat is not None:
            if not (host_heat.is_pinned() and host_heat.dtype == torch.float32 and host_heat.is_contiguous() and
                    host_heat.shape[0] == total):
                raise ValueError("host_heat must be one contiguous pinned float32 tensor covering every frame")
            self.heat = host_heat                                          # stays in host memory
            fill(self.est, "estimated_local_skeleton", every), fill(self.cams, "camera_pose_list", every)
            if has_gt:
                fill(self.gt, "gt_global_skeleton", every)
            copy_stream = None
        else:
            self.heat = alloc("heatmap_list", torch.float32)               # [F,H,W,15] (pickle's HWC layout, as is)
        if host_heat is not None:
            pass
        elif copy_stream is None:
            fill(self.est, "estimated_local_skeleton", every), fill(self.cams, "camera_pose_list", every)
            if has_gt:
                fill(self.gt, "gt_global_skeleton", every)
            fill(self.heat, "heatmap_list", every)
        else:
            main = torch.cuda.current_stream(dev)
            copy_stream.wait_stream(main)                # the buffers may be recycled blocks still in use on `main`
            with torch.cuda.stream(copy_stream):
                # the small per-frame arrays of every clip first: the host-side bookkeeping below needs them all
                fill(self.est, "estimated_local_skeleton", every), fill(self.cams, "camera_pose_list", every)
                if has_gt:
                    fill(self.gt, "gt_global_skeleton", every)
                small = torch.cuda.Event()
                small.record(copy_stream)
                # heat maps in pieces of whole windows, every piece with its completion event
                for i in every:
                    if self.n_windows[i] == 0:
                        continue
                    src = src_of(clips[i], "heatmap_list")
                    for a, _, f0, f1 in upload_pieces(self.n_windows[i], self.n_frames[i], engine.T, OVERLAP,
                                                      min_piece_windows):
                        self.heat[offs[i] + f0:offs[i] + f1].copy_(src[f0:f1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        self.ready_events.append(ev)
                        self.ready_first_windows.append(int(self.window_offsets[i]) + a)
            main.wait_event(small)
            for t in (self.est, self.cams, self.gt, self.heat):
                if t is not None:
                    t.record_stream(copy_stream)
        # mean bone length per clip from the fp32 cast of the whole clip's local estimate
        # (BodyPoseOptimizer.__init__, optimizer.py:42-43, 333, 343); a clip that is one rank's shard of a longer
        # sequence brings the whole sequence's lengths along ("mean_bone_length", computed before sharding)
        est32 = self.est.to(torch.float32)
        bone = torch.linalg.vector_norm(est32 - est32[:, parents, :], dim=-1)   # [F,15]
        def given(i):
            mb = clips[i]["mean_bone_length"]          # (a pinned tensor crosses without synchronising the host)
            mb = mb if isinstance(mb, torch.Tensor) else torch.as_tensor(np.asarray(mb, dtype=np.float32))
            return mb.to(device=dev, dtype=torch.float32, non_blocking=True)
        self.mean_bone = torch.stack([given(i) if "mean_bone_length" in clips[i] else bone[offs[i]:offs[i + 1]].mean(0)
                                      for i in range(len(clips))])

    def replicate(self, R):
        """Every window R times (independent solves from their own noise that share the clips' inputs in HBM):
        the 1e5-window point of BASELINE configs[4] without 200 GB of heat maps.  Replica r of clip i becomes
        clip r * n_clips + i."""
        R = int(R)
        if R <= 1:
            return self
        n_clips = len(self.n_windows)
        self.frame_base, self.clip_idx = self.frame_base.repeat(R), self.clip_idx.repeat(R)
        self.frame_idx = self.frame_idx.repeat(R, 1)
        self.n_windows = list(self.n_windows) * R
        self.starts = list(self.starts) * R
        self.window_offsets = np.concatenate([[0], np.cumsum(self.n_windows)]).astype(np.int64)
        self.W = int(self.window_offsets[-1])
        self.n_clips_unique = n_clips
        return self

    def clip_slices(self, min_windows=96):
        """Slice starts for the library: the upload pieces when the batch was uploaded on a copy stream, else
        the clips; None (automatic slicing) when a start is odd or a slice would be too small."""
        if self.ready_first_windows:
            starts = list(self.ready_first_windows)
        else:
            starts = [int(self.window_offsets[i]) for i, n in enumerate(self.n_windows) if n > 0]
        sizes = np.diff(starts + [self.W])
        if len(starts) < 2 or len(starts) > 64 or any(s % 2 for s in starts) or sizes.min() < min_windows:
            return None
        return starts

    def gather(self, per_frame):
        return per_frame[self.frame_idx]                                    # [W,T,...]


class SequenceOptimizer:
    """Two-stage optimisation of whole clips, all windows in one batch."""

    def __init__(self, engine: Engine, vae_weight=0.0, smoothness_weight=0.001, bone_length_weight=0.01,
                 weight_3d=0.01, reproj_weight=0.01, lr=2, max_iter=25):
        self.engine = engine
        self.w_local, self.w_global = stage_weights(vae_weight, smoothness_weight, bone_length_weight, weight_3d,
                                                    reproj_weight)
        self.params = lbfgs_params(lr=lr, max_iter=max_iter)
        self.slice_by_clip = False      # resident inputs: let the library cut equal slices

    def solve(self, batch: WindowBatch, eps=None, want_trace=False):
        """Runs both stages for every window of ``batch``; returns per-window device tensors."""
        eng = self.engine
        W = batch.W
        eng.set_heat_layout(getattr(batch, "planar", False))
        if eps is None:
            eps = torch.randn(W, 2, eng.n, device=eng.device, dtype=torch.float32)
        eps = eng._dev(eps, torch.float32)
        x_local64 = batch.gather(batch.est)                               # [W,T,15,3] f64
        cams_w = batch.gather(batch.cams)                                 # [W,T,4,4] f64
        if not want_trace:
            if batch.ready_events:
                # inputs still in flight: one library slice per clip, each starting when its clip has landed
                slices = batch.clip_slices()
                eng.set_slices(slices)
                if slices is not None:
                    eng.set_ready_events(batch.ready_first_windows, batch.ready_events)
                else:                                   # cannot slice by clip: wait for everything
                    torch.cuda.current_stream(eng.device).wait_event(batch.ready_events[-1])
            elif self.slice_by_clip:
                eng.set_slices(batch.clip_slices())
            # the whole per-window path in one call: local stage -> SLAM transform -> global stage, slice by
            # slice on the library's internal streams                     optimizer.py:386-419
            sol = eng.solve_windows(x_local64.to(torch.float32), batch.heat, batch.frame_base, batch.clip_idx,
                                    batch.mean_bone, cams_w, eps, self.w_local, self.w_global, self.params)
            sol.update(cams=cams_w, x_local64=x_local64)
            return sol
        # local stage                                                     optimizer.py:386
        loc = eng.solve_stage(0, x_local64.to(torch.float32), batch.heat, batch.frame_base, batch.clip_idx,
                              batch.mean_bone, eps[:, 0], self.w_local, self.params, want_trace=want_trace)
        # SLAM: camera-relative frames of the window -> first camera's frame   optimizer.py:394-398
        rel64, rel32 = eng.relative_global(loc["pose"], cams_w)
        # global stage                                                    optimizer.py:414
        glo = eng.solve_stage(1, rel32, None, None, batch.clip_idx, batch.mean_bone, eps[:, 1], self.w_global,
                              self.params, want_trace=want_trace)
        return dict(local=loc, glob=glo, rel64=rel64, cams=cams_w, x_local64=x_local64)

    def stitch(self, batch: WindowBatch, sol, final_smooth=True):
        """Per clip: the six merged sequences of optimizer.py:442-450 (device float64 tensors)."""
        eng = self.engine
        cams_w = sol["cams"]
        est_rel64, _ = eng.relative_global(sol["x_local64"], cams_w, want_f32=False)
        est_global = eng.to_global(est_rel64, cams_w)                     # optimizer.py:400
        mid_global = eng.to_global(sol["rel64"], cams_w)                  # optimizer.py:401-402
        opt_global = eng.to_global(sol["glob"]["pose"], cams_w)           # optimizer.py:421
        gt_w = batch.gather(batch.gt) if batch.gt is not None else None
        mid_local = sol["local"]["pose"].to(torch.float64)
        out = []
        w0 = 0
        for nw in batch.n_windows:
            sl = slice(w0, w0 + nw)
            w0 += nw
            if nw == 0:
                out.append(None)
                continue
            r = dict(final_estimated_seq=eng.merge_windows(est_global[sl]),
                     mid_estimated_seq=eng.merge_windows(mid_global[sl]),
                     mid_local_pose_seq=eng.merge_windows(mid_local[sl]).to(torch.float32),
                     final_optimized_seq=eng.merge_windows(opt_global[sl]),
                     final_estimated_local_seq=eng.merge_windows(sol["x_local64"][sl]),
                     final_gt_seq=None if gt_w is None else eng.merge_windows(gt_w[sl]))
            if final_smooth:
                r["final_optimized_seq"] = eng.gaussian_smooth(r["final_optimized_seq"], 1.0)
            out.append(r)
        return out

    def run(self, clips, eps=None, final_smooth=True, want_trace=False):
        batch = WindowBatch(self.engine, clips)
        sol = self.solve(batch, eps=eps, want_trace=want_trace)
        return batch, sol, self.stitch(batch, sol, final_smooth=final_smooth)
