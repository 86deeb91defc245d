"""PyTorch-CPU port of the reference's window optimiser, used ONLY as the timed
CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``) and as a
second checker in tests.  Oracle: test infrastructure only (oracle/__init__.py).

It performs the same ATen work per closure evaluation as reference
optimizer.py:226-276 on the same third-party optimiser (``torch.optim.LBFGS``,
strong Wolfe): decoder forward through conv_transpose1d / batch_norm / leaky_relu,
the five energy terms with grid_sample, autograd backward — including the
reference's unused weight gradients (its network parameters keep
requires_grad=True), because that is what the reference's CPU
path costs.  Pass ``weight_grads=False`` to time the leaner variant.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

PARENTS = [0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13]


class TorchVae:
    def __init__(self, state_dict, weight_grads=True, seq_len=10):
        self.p = {}
        for k, v in state_dict.items():
            t = torch.from_numpy(np.array(v))
            if t.is_floating_point() and "running_" not in k:
                t.requires_grad_(weight_grads)
            self.p[k] = t
        self.T = seq_len

    def _bn(self, x, prefix):
        p = self.p
        return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".weight"],
                            p[prefix + ".bias"], training=False, eps=1e-5)

    def encode(self, pose):                                  # SeqConvVAE.py:97-116
        p = self.p
        h = pose.permute(0, 2, 1).contiguous()
        for i in range(5):
            h = F.leaky_relu(self._bn(F.conv1d(h, p[f"encoder.{i}.0.weight"], p[f"encoder.{i}.0.bias"], padding=1),
                                      f"encoder.{i}.1"))
        h = torch.flatten(h, start_dim=1)
        return F.linear(h, p["fc_mu.weight"], p["fc_mu.bias"]), F.linear(h, p["fc_var.weight"], p["fc_var.bias"])

    def decode_to_bodypose(self, z):                         # SeqConvVAE.py:131-140
        p = self.p
        h = F.linear(z, p["decoder_input.weight"], p["decoder_input.bias"]).view(-1, 512, self.T)
        for i in range(4):
            h = F.leaky_relu(self._bn(F.conv_transpose1d(h, p[f"decoder.{i}.0.weight"], p[f"decoder.{i}.0.bias"],
                                                         padding=1), f"decoder.{i}.1"))
        h = F.leaky_relu(self._bn(F.conv_transpose1d(h, p["final_layer.0.weight"], p["final_layer.0.bias"],
                                                     padding=1), "final_layer.1"))
        h = F.conv1d(h, p["final_layer.3.weight"], p["final_layer.3.bias"], padding=1)
        return h.permute(0, 2, 1).view(-1, self.T, 15, 3)


class TorchPortOptimizer:
    """One stage solver (reference BodyPoseOptimizer)."""

    def __init__(self, state_dict, camera, mean_bone, weights, max_iter=25, lr=2, weight_grads=True):
        self.vae = TorchVae(state_dict, weight_grads)
        poly, cx, cy = camera
        self.poly = [float(c) for c in poly]
        self.cx, self.cy = float(cx), float(cy)
        self.mean_bone = torch.as_tensor(np.asarray(mean_bone), dtype=torch.float32)
        self.w3d, self.ws, self.wb, self.wv, self.wr = [float(w) for w in weights]
        self.max_iter, self.lr = max_iter, lr

    def _project(self, p3):                                  # FishEyeCalibrated.py:96-129
        x, y, z = p3[:, 0], p3[:, 1], -p3[:, 2]
        norm = torch.sqrt(x * x + y * y)
        if not bool((norm != 0).all()):
            raise Exception("norm is zero!")
        theta = torch.atan(z / norm)
        rho = self.poly[0]
        t_i = 1.0
        for c in self.poly[1:]:
            t_i = t_i * theta
            rho = rho + t_i * c
        inv = 1.0 / norm
        return torch.stack([x * inv * rho + self.cx, y * inv * rho + self.cy], dim=1)

    def total_loss(self, z, x0, maps):                       # optimizer.py:226-240
        x = self.vae.decode_to_bodypose(z).squeeze(0).contiguous()
        e3d = torch.sum(torch.square(x - x0))
        v = x[:-1] - x[1:]
        a = v[:-1] - v[1:]
        esm = torch.sum(torch.square(a))
        bone = torch.norm(x - x[:, PARENTS, :], dim=-1)
        eb = torch.sum(torch.square(bone - self.mean_bone))
        ev = torch.sum(torch.square(x))
        if self.wr == 0:
            er = 0
        else:
            uv = self._project(x.view(-1, 3))
            g = torch.stack([(uv[:, 0] - 128 - 512) / 512, (uv[:, 1] - 512) / 512], dim=1).view(-1, 1, 1, 2)
            er = -torch.sum(F.grid_sample(maps, g, align_corners=True))
        return self.w3d * e3d + self.ws * esm + self.wb * eb + self.wv * ev + self.wr * er

    def solve(self, pose_in, heat_hwc, eps, trace=None):     # optimizer.py:242-276
        x0 = torch.from_numpy(np.asarray(pose_in)).float()
        maps = None
        if self.wr != 0:
            hs = torch.from_numpy(np.asarray(heat_hwc)).float().permute(0, 3, 1, 2).contiguous()
            maps = hs.view(-1, 1, hs.shape[-2], hs.shape[-1])
        with torch.no_grad():
            mu, logvar = self.vae.encode(x0.view(1, self.vae.T, 45))
            z0 = torch.from_numpy(np.asarray(eps, dtype=np.float32)).view(1, -1) * torch.exp(0.5 * logvar) + mu
        z = torch.nn.Parameter(z0.detach().clone())
        opt = torch.optim.LBFGS([z], lr=self.lr, max_iter=self.max_iter, tolerance_change=1e-6,
                                line_search_fn="strong_wolfe")

        def closure():
            opt.zero_grad()
            loss = self.total_loss(z, x0, maps)
            loss.backward()
            if trace is not None:
                trace.append((float(loss.detach()), z.detach().clone().numpy().reshape(-1)))
            return loss

        opt.step(closure)
        with torch.no_grad():
            out = self.vae.decode_to_bodypose(z).squeeze(0).numpy().copy()
        return out, dict(n_iter=opt.state[z]["n_iter"], func_evals=opt.state[z]["func_evals"])
