"""numpy restatement of the motion VAE (reference networks/models/SeqConvVAE.py)
in inference mode: encoder -> (mu, std), decoder -> body pose, and the
decoder's vector-Jacobian product w.r.t. the latent (what autograd computes at
reference optimizer.py:267; weight gradients are never used).

Oracle: test infrastructure only (see oracle/__init__.py).  Layers are applied
one by one exactly as the reference's nn.Sequential does (no folding/fusion),
so this file doubles as the specification the fused CUDA path is checked
against.
"""
from __future__ import annotations

import numpy as np

EPS_BN = 1e-5
SLOPE = 0.01            # nn.LeakyReLU() default negative_slope


def _shift(x, s):
    """y[..., t] = x[..., t+s], zero outside."""
    y = np.zeros_like(x)
    T = x.shape[-1]
    if s == 0:
        y[...] = x
    elif s > 0:
        y[..., :T - s] = x[..., s:]
    else:
        y[..., -s:] = x[..., :T + s]
    return y


def conv1d_k3(x, w, b):
    """nn.Conv1d(k=3, s=1, p=1): y[o,t] = b[o] + sum_i sum_k x[i,t+k-1] w[o,i,k].  x (B,Cin,T)."""
    y = sum(np.einsum("oi,bit->bot", w[:, :, k], _shift(x, k - 1)) for k in range(3))
    return y + b[None, :, None]


def conv1d_k3_bwd(dy, w):
    """dL/dx of conv1d_k3: dx[i,t] = sum_o sum_k dy[o,t-k+1] w[o,i,k]."""
    return sum(np.einsum("oi,bot->bit", w[:, :, k], _shift(dy, 1 - k)) for k in range(3))


def convT1d_k3(x, w, b):
    """nn.ConvTranspose1d(k=3, s=1, p=1), weight (in,out,k):
    y[o,t] = b[o] + sum_i sum_k x[i,t+1-k] w[i,o,k]  (SURVEY.md A.7)."""
    y = sum(np.einsum("io,bit->bot", w[:, :, k], _shift(x, 1 - k)) for k in range(3))
    return y + b[None, :, None]


def convT1d_k3_bwd(dy, w):
    """dx[i,t] = sum_o sum_k dy[o,t-1+k] w[i,o,k]."""
    return sum(np.einsum("io,bot->bit", w[:, :, k], _shift(dy, k - 1)) for k in range(3))


def bn_eval(x, sd, prefix):
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    scale = g / np.sqrt(v + x.dtype.type(EPS_BN))
    return (x - m[None, :, None]) * scale[None, :, None] + b[None, :, None], scale


def lrelu(x):
    return np.where(x > 0, x, x * x.dtype.type(SLOPE))


def lrelu_grad(pre):
    return np.where(pre > 0, pre.dtype.type(1), pre.dtype.type(SLOPE))


def _cast(sd, dtype):
    return {k: (np.asarray(v).astype(dtype) if np.asarray(v).dtype.kind == "f" else np.asarray(v))
            for k, v in sd.items()}


class VaeNp:
    def __init__(self, state_dict, dtype=np.float32, seq_len=10):
        self.sd = _cast(state_dict, dtype)
        self.dtype = np.dtype(dtype)
        self.T = seq_len

    # ---- encoder: SeqConvVAE.py:97-116, 184-189
    def encode(self, pose):
        """pose (B,T,45) -> mu, std (B,latent)."""
        sd = self.sd
        h = np.transpose(pose.astype(self.dtype), (0, 2, 1))            # permute((0,2,1))
        for i in range(5):
            h = conv1d_k3(h, sd[f"encoder.{i}.0.weight"], sd[f"encoder.{i}.0.bias"])
            h, _ = bn_eval(h, sd, f"encoder.{i}.1")
            h = lrelu(h)
        flat = h.reshape(h.shape[0], -1)                                 # torch.flatten: index c*T+t
        mu = flat @ sd["fc_mu.weight"].T + sd["fc_mu.bias"]
        logvar = flat @ sd["fc_var.weight"].T + sd["fc_var.bias"]
        return mu, np.exp(self.dtype.type(0.5) * logvar)

    def latent(self, pose, eps):
        """get_latent_space + reparameterize with injected noise: z0 = eps*std + mu."""
        mu, std = self.encode(pose)
        return eps.astype(self.dtype) * std + mu

    # ---- decoder: SeqConvVAE.py:131-140
    def decode(self, z, keep=False):
        """z (B,latent) -> pose (B,T,15,3); with keep=True also the saved
        pre-activations needed by decode_vjp."""
        sd = self.sd
        B = z.shape[0]
        h = z.astype(self.dtype) @ sd["decoder_input.weight"].T + sd["decoder_input.bias"]
        h = h.reshape(B, 512, self.T)                                    # view(-1, 512, T)
        saved = []
        for i in range(4):
            h = convT1d_k3(h, sd[f"decoder.{i}.0.weight"], sd[f"decoder.{i}.0.bias"])
            h, scale = bn_eval(h, sd, f"decoder.{i}.1")
            saved.append((h, scale))
            h = lrelu(h)
        h = convT1d_k3(h, sd["final_layer.0.weight"], sd["final_layer.0.bias"])
        h, scale = bn_eval(h, sd, "final_layer.1")
        saved.append((h, scale))
        h = lrelu(h)
        out = conv1d_k3(h, sd["final_layer.3.weight"], sd["final_layer.3.bias"])   # (B,45,T)
        pose = np.transpose(out, (0, 2, 1)).reshape(B, self.T, 15, 3)   # x[t,j,k] = out[3j+k,t]
        return (pose, saved) if keep else pose

    def decode_vjp(self, saved, dpose):
        """dL/dz given dL/dpose (B,T,15,3) and the activations of decode(keep=True)."""
        sd = self.sd
        B = dpose.shape[0]
        d = np.transpose(dpose.reshape(B, self.T, 45).astype(self.dtype), (0, 2, 1))
        d = conv1d_k3_bwd(d, sd["final_layer.3.weight"])
        pre, scale = saved[4]
        d = d * lrelu_grad(pre) * scale[None, :, None]
        d = convT1d_k3_bwd(d, sd["final_layer.0.weight"])
        for i in (3, 2, 1, 0):
            pre, scale = saved[i]
            d = d * lrelu_grad(pre) * scale[None, :, None]
            d = convT1d_k3_bwd(d, sd[f"decoder.{i}.0.weight"])
        return d.reshape(B, -1) @ sd["decoder_input.weight"]
