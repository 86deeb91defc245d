"""Thin host-side owner of a gem_ctx: device memory and streams come from
PyTorch, all arithmetic happens in libgem_b200.so (include/gem_b200.h)."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np
import torch

from . import _lib
from ._lib import EnergyWeights, GemError, LbfgsParams, Layer, VaeWeights, check
from .vae_prep import PreparedVae

KINEMATIC_PARENTS = (0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13)   # reference optimizer.py:34


def lbfgs_params(lr=2, max_iter=25, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-6) -> LbfgsParams:
    """torch.optim.LBFGS defaults as the reference constructs it (optimizer.py:261-262)."""
    if max_eval is None:
        max_eval = max_iter * 5 // 4
    return LbfgsParams(float(lr), int(max_iter), int(max_eval), float(tolerance_grad), float(tolerance_change))


def energy_weights(weight_3d, smooth_weight, bone_length_weight, vae_weight, reproj_weight) -> EnergyWeights:
    return EnergyWeights(float(weight_3d), float(smooth_weight), float(bone_length_weight), float(vae_weight),
                         float(reproj_weight))


class _DevPtr:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def heat_layout_code(layout) -> int:
    """0 (HWC, the pickle's), 1 (planar) or 2 (tiled) from a bool, an int or a name."""
    if isinstance(layout, str):
        return {"hwc": 0, "planar": 1, "tiled": 2}[layout]
    code = int(layout)
    if code not in (0, 1, 2):
        raise ValueError("heat layout must be 0 / 'hwc', 1 / 'planar' or 2 / 'tiled'")
    return code


def heat_dims(shape, layout):
    """(H, W, J) of heat maps whose per-frame shape ends `shape`, in the given layout."""
    code = heat_layout_code(layout)
    if code == 0:
        return int(shape[-3]), int(shape[-2]), int(shape[-1])
    if code == 1:
        return int(shape[-2]), int(shape[-1]), int(shape[-3])
    return int(shape[-4]) * 4, int(shape[-3]) * 8, int(shape[-5])


def tile_heat(planar_maps):
    """[..., H, W] maps -> [..., H/4, W/8, 4, 8] (the tiled layout's per-map order), a strided view."""
    *lead, H, W = planar_maps.shape
    if H % 4 or W % 8:
        raise ValueError("tiled heat maps need H % 4 == 0 and W % 8 == 0")
    n = len(lead)
    v = planar_maps.reshape(*lead, H // 4, 4, W // 8, 8)
    return v.permute(*range(n), n, n + 2, n + 1, n + 3)


def untile_heat(tiled_maps):
    """[..., H/4, W/8, 4, 8] -> [..., H, W] (a copy)."""
    *lead, th, tw, r, c = tiled_maps.shape
    n = len(lead)
    return tiled_maps.permute(*range(n), n, n + 2, n + 1, n + 3).reshape(*lead, th * r, tw * c)


class Engine:
    def __init__(self, max_windows, device=None, latent_dim=2048, seq_len=10, num_joints=15, heat_hw=(64, 64),
                 max_history=24):
        if not torch.cuda.is_available():
            raise GemError("globalegomocap_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        self.max_windows, self.n, self.T, self.J = int(max_windows), latent_dim, seq_len, num_joints
        self.H, self.Wd = heat_hw
        self.max_history = max_history
        self._ctx = C.c_void_p()
        check(self.lib.gem_ctx_create(C.byref(self._ctx), self.device.index, self.max_windows, latent_dim, seq_len,
                                      num_joints, self.H, self.Wd, max_history))
        self._vae = {}
        self._keep = []
        self.heat_planar = 0            # heat-map layout code (heat_layout_code)

    # ------------------------------------------------------------------ housekeeping
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.gem_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def scratch_bytes(self):
        return int(self.lib.gem_ctx_scratch_bytes(self._ctx))

    def launch_count(self) -> int:
        return int(self.lib.gem_ctx_launch_count(self._ctx))

    def set_profiling(self, enable: bool):
        check(self.lib.gem_ctx_set_profiling(self._ctx, int(bool(enable))))

    def read_profile(self):
        """{tag: (launches, total_ms)} of the event pairs recorded since the last read (synchronises)."""
        n_max = 64
        tags, counts = (C.c_int32 * n_max)(), (C.c_int32 * n_max)()
        ms, n = (C.c_float * n_max)(), C.c_int32(0)
        check(self.lib.gem_ctx_read_profile(self._ctx, n_max, tags, counts, ms, C.byref(n)))
        return {int(tags[i]): (int(counts[i]), float(ms[i])) for i in range(n.value)}

    def set_gemm_mode(self, mode: int):
        check(self.lib.gem_ctx_set_gemm_mode(self._ctx, int(mode)))

    def set_chunks(self, n_chunks: int):
        """Number of window slices a stage runs concurrently on internal streams (results are unaffected)."""
        check(self.lib.gem_ctx_set_chunks(self._ctx, int(n_chunks)))

    def set_slices(self, first_windows):
        """Explicit slice boundaries (first window of each slice, even, starting at 0); None/[] = automatic."""
        fw = list(first_windows or [])
        arr = (C.c_int32 * max(len(fw), 1))(*fw)
        check(self.lib.gem_ctx_set_slices(self._ctx, len(fw), arr))

    def set_ready_events(self, first_windows, events):
        """One-shot for the next solve: ``events[i]`` (torch.cuda.Event) completes when the inputs of the windows
        from ``first_windows[i]`` up to the next entry are resident."""
        n = len(events)
        fw = (C.c_int32 * max(n, 1))(*[int(v) for v in first_windows])
        ev = (C.c_void_p * max(n, 1))(*[C.c_void_p(e.cuda_event) for e in events])
        self._keep_events = list(events)
        check(self.lib.gem_ctx_set_ready_events(self._ctx, n, fw, ev))

    def set_texel_cache(self, mode: int):
        """-1 auto (on when the heat maps are pinned host memory), 0 off, 1 on; results are unaffected."""
        check(self.lib.gem_ctx_set_texel_cache(self._ctx, int(mode)))

    def set_heat_layout(self, planar):
        """False / 0 / "hwc": heat maps are [frames, H, W, J] (the pickle's layout); True / 1 / "planar": [frames, J, H, W];
        2 / "tiled": [frames, J, H/4, W/8, 4, 8], every map as tiles of 4 rows x 8 texels (one 128-byte line each: what the
        zero-copy path fetches per PCIe request) (include/gem_b200.h: gem_ctx_set_heat_layout).  Results are
        bit-identical in all layouts."""
        code = heat_layout_code(planar)
        if code != self.heat_planar:
            check(self.lib.gem_ctx_set_heat_layout(self._ctx, code))
            self.heat_planar = code

    def texel_cache_stats(self, enable: bool):
        """(lookups, texels fetched from the map) counted since the previous call; switches the counting on or off."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(self.lib.gem_ctx_texel_cache_stats(self._ctx, int(bool(enable)), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def _heat(self, heat):
        """Heat maps for the solver: a CUDA tensor, or — left where it is — a pinned host tensor that the energy
        kernel reads over PCIe through its texel cache (include/gem_b200.h: gem_ctx_set_texel_cache)."""
        if heat is None:
            return None
        if (isinstance(heat, torch.Tensor) and not heat.is_cuda and heat.dtype == torch.float32 and
                heat.is_contiguous() and heat.is_pinned()):
            return heat
        return self._dev(heat, torch.float32)

    def _check_heat(self, heat, frame_base):
        """The C ABI takes a raw map pointer and indexes it with the ctx's (H, Wd, J): a pickle with another map
        resolution / joint count, or a window that runs past the last frame, must not reach the kernel."""
        if heat is None:
            return
        want = {0: (self.H, self.Wd, self.J), 1: (self.J, self.H, self.Wd),
                2: (self.J, self.H // 4, self.Wd // 8, 4, 8)}[int(self.heat_planar)]
        if tuple(heat.shape[-len(want):]) != want:
            raise GemError(f"heat maps are {tuple(heat.shape[-len(want):])}, the engine expects {want} "
                           f"({('H, W, joints', 'joints, H, W: planar layout', 'joints, H/4, W/8, 4, 8: tiled layout')[int(self.heat_planar)]})")
        if frame_base is None or (isinstance(frame_base, torch.Tensor) and frame_base.is_cuda):
            return          # device-resident indices were range-checked by whoever built them (WindowBatch does)
        fb = np.asarray(frame_base)
        if fb.size:
            n_frames = int(np.prod(heat.shape)) // (self.H * self.Wd * self.J)
            lo, hi = int(fb.min()), int(fb.max())
            if lo < 0 or hi + self.T > n_frames:
                raise GemError(f"window frame range [{lo}, {hi + self.T}) leaves the {n_frames} heat-map frames")

    def _dev(self, t, dtype):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        t = t.to(device=self.device, dtype=dtype)
        return t if t.is_contiguous() else t.contiguous()

    @staticmethod
    def _p(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def _check_w(self, W):
        if W > self.max_windows:
            raise GemError(f"{W} windows exceed the engine capacity {self.max_windows}")

    # ------------------------------------------------------------------ configuration
    def set_camera(self, poly, cx, cy):
        arr = (C.c_double * len(poly))(*[float(c) for c in poly])
        check(self.lib.gem_ctx_set_camera(self._ctx, arr, len(poly), float(cx), float(cy)))

    def set_camera_json(self, path):
        """Reads the reference's calibration JSON (FishEyeCalibrated.py:8-14)."""
        with open(path) as f:
            cal = json.load(f)
        intr = cal["intrinsic"]
        self.set_camera(cal["polynomialW2C"], intr[0][2], intr[1][2])

    def set_vae(self, which: int, state_dict):
        prep = state_dict if isinstance(state_dict, PreparedVae) else PreparedVae(
            state_dict, self.device, seq_len=self.T, latent_dim=self.n, channels=self.J * 3)
        if self._vae.get(which) is prep:          # the same prepared weights are already in the ctx
            return prep
        vw = VaeWeights()
        for name in ("dec", "dec_bwd", "enc"):
            arr = getattr(vw, name)
            for i, l in enumerate(getattr(prep, name)):
                arr[i] = Layer(l["w"].data_ptr(), 0 if l["bias"] is None else l["bias"].data_ptr(), l["taps"], l["k"],
                               l["n"])
        check(self.lib.gem_ctx_set_vae(self._ctx, int(which), C.byref(vw)))
        self._vae[which] = prep
        return prep

    # ------------------------------------------------------------------ fused energy + gradient
    def energy_grad(self, pose, pose0, heat, frame_base, clip, mean_bone, weights: EnergyWeights, want_terms=True):
        pose, pose0 = self._dev(pose, torch.float32), self._dev(pose0, torch.float32)
        W = pose.shape[0]
        self._check_w(W)
        heat = None if heat is None else self._dev(heat, torch.float32)
        self._check_heat(heat, frame_base)
        frame_base = None if frame_base is None else self._dev(frame_base, torch.int64)
        clip = self._dev(clip, torch.int32)
        mean_bone = self._dev(mean_bone, torch.float32).reshape(-1, self.J)
        E = torch.empty(W, dtype=torch.float32, device=self.device)
        terms = torch.empty(W, 5, dtype=torch.float32, device=self.device) if want_terms else None
        grad = torch.empty_like(pose)
        status = torch.zeros(W, dtype=torch.int32, device=self.device)
        check(self.lib.gem_energy_grad(self._ctx, self.stream, W, self._p(pose), self._p(pose0), self._p(heat),
                                       self._p(frame_base), self._p(clip), self._p(mean_bone), C.byref(weights),
                                       self._p(E), self._p(terms), self._p(grad), self._p(status)))
        return E, terms, grad, status

    def gemm(self, a, b, bias=None, leaky_relu=False, tensor_cores=False):
        """c = act(bias + a @ b) on the library's GEMM kernels (test hook)."""
        a, b = self._dev(a, torch.float32), self._dev(b, torch.float32)
        bias = None if bias is None else self._dev(bias, torch.float32)
        M, K = a.shape
        N = b.shape[1]
        c = torch.empty(M, N, dtype=torch.float32, device=self.device)
        check(self.lib.gem_gemm(self._ctx, self.stream, M, N, K, self._p(a), K, self._p(b), self._p(bias),
                                int(leaky_relu), self._p(c), N, int(tensor_cores)))
        return c

    # ------------------------------------------------------------------ VAE pieces
    def decode(self, which, z):
        z = self._dev(z, torch.float32)
        W = z.shape[0]
        self._check_w(W)
        pose = torch.empty(W, self.T, self.J, 3, dtype=torch.float32, device=self.device)
        check(self.lib.gem_decode(self._ctx, self.stream, int(which), W, self._p(z), self._p(pose)))
        return pose

    def decode_vjp(self, which, dpose):
        dpose = self._dev(dpose, torch.float32)
        W = dpose.shape[0]
        dz = torch.empty(W, self.n, dtype=torch.float32, device=self.device)
        check(self.lib.gem_decode_vjp(self._ctx, self.stream, int(which), W, self._p(dpose), self._p(dz)))
        return dz

    def encode(self, which, pose, eps):
        pose, eps = self._dev(pose, torch.float32), self._dev(eps, torch.float32)
        W = pose.shape[0]
        self._check_w(W)
        z0 = torch.empty(W, self.n, dtype=torch.float32, device=self.device)
        mu, std = torch.empty_like(z0), torch.empty_like(z0)
        check(self.lib.gem_encode(self._ctx, self.stream, int(which), W, self._p(pose), self._p(eps), self._p(z0),
                                  self._p(mu), self._p(std)))
        return z0, mu, std

    # ------------------------------------------------------------------ batched L-BFGS driven from the host
    def lbfgs_begin(self, z0, params: LbfgsParams):
        z0 = self._dev(z0, torch.float32)
        self._lb_W = z0.shape[0]
        self._check_w(self._lb_W)
        check(self.lib.gem_lbfgs_begin(self._ctx, self.stream, self._lb_W, self._p(z0), C.byref(params)))

    def lbfgs_trial(self):
        ptr = self.lib.gem_lbfgs_trial(self._ctx)
        return torch.as_tensor(_DevPtr(ptr, (self._lb_W, self.n)), device=self.device)

    def lbfgs_x(self):
        ptr = self.lib.gem_lbfgs_x(self._ctx)
        return torch.as_tensor(_DevPtr(ptr, (self._lb_W, self.n)), device=self.device)

    def lbfgs_advance(self, loss, grad):
        loss, grad = self._dev(loss, torch.float32), self._dev(grad, torch.float32)
        check(self.lib.gem_lbfgs_advance(self._ctx, self.stream, self._lb_W, self._p(loss), self._p(grad)))

    def lbfgs_stats(self):
        W = self._lb_W
        n_iter = torch.empty(W, dtype=torch.int32, device=self.device)
        evals, fin = torch.empty_like(n_iter), torch.empty_like(n_iter)
        t = torch.empty(W, dtype=torch.float64, device=self.device)
        loss = torch.empty_like(t)
        check(self.lib.gem_lbfgs_stats(self._ctx, self.stream, W, self._p(n_iter), self._p(evals), self._p(fin),
                                       self._p(t), self._p(loss), None, 0))
        return dict(n_iter=n_iter, func_evals=evals, finished=fin, t=t, loss=loss)

    # ------------------------------------------------------------------ one optimisation stage for W windows
    def solve_stage(self, which, pose0, heat, frame_base, clip, mean_bone, eps, weights: EnergyWeights,
                    params: LbfgsParams, want_trace=False, out=None):
        pose0, eps = self._dev(pose0, torch.float32), self._dev(eps, torch.float32)
        W = pose0.shape[0]
        self._check_w(W)
        heat = self._heat(heat)
        self._check_heat(heat, frame_base)
        frame_base = None if frame_base is None else self._dev(frame_base, torch.int64)
        clip = self._dev(clip, torch.int32)
        mean_bone = self._dev(mean_bone, torch.float32).reshape(-1, self.J)
        pose_out = out if out is not None else torch.empty(W, self.T, self.J, 3, dtype=torch.float32,
                                                           device=self.device)
        trace = (torch.empty(W, params.max_eval + 1, dtype=torch.float32, device=self.device)
                 if want_trace else None)
        n_iter = torch.empty(W, dtype=torch.int32, device=self.device)
        evals = torch.empty_like(n_iter)
        status = torch.empty(W, dtype=torch.int32, device=self.device)
        check(self.lib.gem_solve_stage(self._ctx, self.stream, int(which), W, self._p(pose0), self._p(heat),
                                       self._p(frame_base), self._p(clip), self._p(mean_bone), self._p(eps),
                                       C.byref(weights), C.byref(params), self._p(pose_out), self._p(trace),
                                       self._p(n_iter), self._p(evals), self._p(status)))
        return dict(pose=pose_out, trace=trace, n_iter=n_iter, func_evals=evals, status=status)

    def solve_windows(self, pose0, heat, frame_base, clip, mean_bone, cams, eps, w_local: EnergyWeights,
                      w_global: EnergyWeights, params: LbfgsParams):
        """Local stage -> SLAM transform -> global stage for W windows, slice by slice (gem_solve_windows)."""
        pose0, eps = self._dev(pose0, torch.float32), self._dev(eps, torch.float32)
        W = pose0.shape[0]
        self._check_w(W)
        heat = self._heat(heat)
        self._check_heat(heat, frame_base)
        frame_base = None if frame_base is None else self._dev(frame_base, torch.int64)
        clip = self._dev(clip, torch.int32)
        mean_bone = self._dev(mean_bone, torch.float32).reshape(-1, self.J)
        cams = self._dev(cams, torch.float64)
        shape = (W, self.T, self.J, 3)
        local = torch.empty(shape, dtype=torch.float32, device=self.device)
        glob = torch.empty(shape, dtype=torch.float32, device=self.device)
        rel32 = torch.empty(shape, dtype=torch.float32, device=self.device)
        rel64 = torch.empty(shape, dtype=torch.float64, device=self.device)
        n_iter = torch.empty(2, W, dtype=torch.int32, device=self.device)
        evals = torch.empty_like(n_iter)
        status = torch.empty(W, dtype=torch.int32, device=self.device)
        check(self.lib.gem_solve_windows(self._ctx, self.stream, W, self._p(pose0), self._p(heat), self._p(frame_base),
                                         self._p(clip), self._p(mean_bone), self._p(cams), self._p(eps),
                                         C.byref(w_local), C.byref(w_global), C.byref(params), self._p(local),
                                         self._p(rel64), self._p(rel32), self._p(glob), self._p(n_iter), self._p(evals),
                                         self._p(status)))
        return dict(local=dict(pose=local, n_iter=n_iter[0], func_evals=evals[0], status=status, trace=None),
                    glob=dict(pose=glob, n_iter=n_iter[1], func_evals=evals[1], trace=None), rel64=rel64, rel32=rel32)

    # ------------------------------------------------------------------ SLAM transforms and stitching
    def relative_global(self, pose, cams, want_f32=True):
        is64 = pose.dtype == torch.float64
        pose = self._dev(pose, torch.float64 if is64 else torch.float32)
        cams = self._dev(cams, torch.float64)
        W = pose.shape[0]
        self._check_w(W)
        out64 = torch.empty(pose.shape, dtype=torch.float64, device=self.device)
        out32 = torch.empty(pose.shape, dtype=torch.float32, device=self.device) if want_f32 else None
        check(self.lib.gem_relative_global(self._ctx, self.stream, W, self._p(pose), int(is64), self._p(cams),
                                           self._p(out64), self._p(out32)))
        return out64, out32

    def to_global(self, pose, cams):
        is64 = pose.dtype == torch.float64
        pose = self._dev(pose, torch.float64 if is64 else torch.float32)
        cams = self._dev(cams, torch.float64)
        W = pose.shape[0]
        self._check_w(W)
        out64 = torch.empty(pose.shape, dtype=torch.float64, device=self.device)
        check(self.lib.gem_to_global(self._ctx, self.stream, W, self._p(pose), int(is64), self._p(cams),
                                     self._p(out64)))
        return out64

    def merge_windows(self, windows, overlap=2):
        windows = self._dev(windows, torch.float64)
        W = windows.shape[0]
        n_frames = (self.T - overlap) * W + overlap
        out = torch.empty(n_frames, self.J, 3, dtype=torch.float64, device=self.device)
        check(self.lib.gem_merge_windows(self._ctx, self.stream, W, int(overlap), self._p(windows), self._p(out)))
        return out

    def gaussian_smooth(self, seq, sigma=1.0):
        seq = self._dev(seq, torch.float64)
        N = seq.shape[0]
        row = seq.numel() // max(N, 1)
        out = torch.empty_like(seq)
        check(self.lib.gem_gaussian_smooth(self._ctx, self.stream, N, row, float(sigma), self._p(seq), self._p(out)))
        return out
