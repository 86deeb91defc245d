"""Runs the UNMODIFIED reference's window optimiser (reference optimizer.py:36-71, 242-276, 386-419) for
`bench.py --impl reference` and bench's `cpu_baseline`, from baseline/_ref (installed by
baseline/install_reference.py; travels to the GPU box) or, in the build container, /root/reference.

Recipe of SURVEY.md Appendix C: `open3d` / `natsort` (absent, unused on the hot path) are stubbed in sys.modules, the
process runs from a scratch CWD that holds random-init checkpoints under the reference's hard-coded
`networks/logs/...` layout, and ConvVAE.reparameterize's noise draw is replaced by a queue so that runs are repeatable.
Nothing of the reference is modified; its own classes and functions do all the work.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")
LOCAL_CKPT = "networks/logs/only_local_full_dataset_latent_2048_len_10_kl_0.5_2/checkpoints/19.pth.tar"
GLOBAL_CKPT = "networks/logs/real_full_dataset_latent_2048_len_10_slide_window_step_1_kl_0.5/checkpoints/19.pth.tar"


def reference_root():
    for root in CANDIDATES:
        if os.path.exists(os.path.join(root, "optimizer.py")):
            return root
    return None


class ReferenceRunner:
    """The reference's two BodyPoseOptimizer instances (local / global stage) as `optimizer.main` builds them
    (optimizer.py:332-358), on CPU."""

    def __init__(self, root, clip, weights, camera_json, max_iter=25):
        import torch
        self.torch = torch            # (the reference runs on cuda when torch sees one, optimizer.py:39: callers that want
                                      #  its CPU path hide the GPUs with CUDA_VISIBLE_DEVICES before importing torch)
        self.root = root
        self.scratch = tempfile.mkdtemp(prefix="gem_ref_run_")
        from globalegomocap_b200 import synthetic as syn
        syn.save_checkpoint(weights[0], os.path.join(self.scratch, LOCAL_CKPT))
        syn.save_checkpoint(weights[1], os.path.join(self.scratch, GLOBAL_CKPT))
        for name in ("open3d", "natsort"):
            if name not in sys.modules:
                m = types.ModuleType(name)
                if name == "natsort":
                    m.natsorted = sorted
                sys.modules[name] = m
        sys.path[:0] = [root, os.path.join(root, "networks")]
        self._cwd = os.getcwd()
        os.chdir(self.scratch)
        import optimizer as ref_opt                                  # the reference's module
        from networks.models.SeqConvVAE import ConvVAE
        self.ref = ref_opt
        self.queue = []
        runner = self

        def reparameterize(self_, mu, logvar):                       # SeqConvVAE.py:159-169 with injected noise
            std = torch.exp(0.5 * logvar)
            return runner.queue.pop(0) * std + mu

        self._orig = ConvVAE.reparameterize
        self._cls = ConvVAE
        ConvVAE.reparameterize = reparameterize
        est = torch.from_numpy(np.asarray(clip["estimated_local_skeleton"])).float()
        common = dict(camera_model_path=camera_json, mean_skeleton=est, latent_dim=2048, network_seq_len=10, seq_len=10,
                      windows_size=1, overlap_size=2, lr=2, max_iter=max_iter)
        self.local = ref_opt.BodyPoseOptimizer(vae_path=LOCAL_CKPT, **common)
        self.local.set_weights(vae_weight=0.0, gmm_weight=0.0, smooth_weight=0.001 / 100, bone_length_weight=0.01,
                               weight_3d=0.01 / 10000, reproj_weight=0.01)                     # optimizer.py:352-354
        self.glob = ref_opt.BodyPoseOptimizer(vae_path=GLOBAL_CKPT, **common)
        self.glob.set_weights(vae_weight=0.0, gmm_weight=0.0, smooth_weight=0.001, bone_length_weight=0.01,
                              weight_3d=0.01, reproj_weight=0)                                  # optimizer.py:356-358
        self.clip = clip

    def close(self):
        self._cls.reparameterize = self._orig
        os.chdir(self._cwd)

    def solve_stage(self, which, x0, heat, eps):
        """One optimize_pose_seq_pytorch_LBFGS call of the local (0) or global (1) optimiser with the energies of every
        closure evaluation recorded: (pose (10,15,3) float32, [E_0, E_1, ...])."""
        torch = self.torch
        opt = self.local if which == 0 else self.glob
        energies = []
        cls = type(opt)
        orig = cls.total_loss

        def traced(self_, hidden):
            e = orig(self_, hidden)
            energies.append(float(e.detach()))
            return e

        cls.total_loss = traced
        try:
            self.queue.append(torch.from_numpy(np.asarray(eps, np.float32)).view(1, -1).to(opt.device))
            x0 = np.asarray(x0)
            pose = opt.optimize_pose_seq_pytorch_LBFGS(x0, np.asarray(heat), x0.copy())
        finally:
            cls.total_loss = orig
        return pose, energies

    def window(self, start, eps):
        """One iteration of the reference's window loop (optimizer.py:372-419): local stage, SLAM transform, global stage."""
        torch = self.torch
        est = np.asarray(self.clip["estimated_local_skeleton"])[start:start + 10]
        heat = np.asarray(self.clip["heatmap_list"])[start:start + 10]
        cams = np.asarray(self.clip["camera_pose_list"])[start:start + 10]
        self.queue.append(torch.from_numpy(np.asarray(eps[0], np.float32)).view(1, -1).to(self.local.device))
        local = self.local.optimize_pose_seq_pytorch_LBFGS(est, heat, est.copy())
        rel = self.ref.get_relative_global_pose_with_camera_matrix(local, cams)
        self.queue.append(torch.from_numpy(np.asarray(eps[1], np.float32)).view(1, -1).to(self.glob.device))
        return self.glob.optimize_pose_seq_pytorch_LBFGS(rel, heat, rel.copy())

    def time_windows(self, first_window, n_windows, seed=0):
        """Seconds per window for n_windows consecutive windows starting at first_window (stride 8)."""
        gen = self.torch.Generator().manual_seed(seed)
        times = []
        for i in range(n_windows):
            eps = self.torch.randn(2, 2048, generator=gen).numpy()
            t0 = time.perf_counter()
            self.window(8 * (first_window + i), eps)
            times.append(time.perf_counter() - t0)
        return times
