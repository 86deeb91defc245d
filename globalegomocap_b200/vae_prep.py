"""Turns a ConvVAE checkpoint (reference networks/models/SeqConvVAE.py:27-92,
loaded as at reference optimizer.py:59-60) into the token-major "tap GEMM"
operands the CUDA kernels consume (include/gem_b200.h: gem_layer).

All algebra is done once, in float64, then rounded to fp32:

* BatchNorm1d in eval mode is affine (running stats, eps 1e-5) and is folded into
  the preceding convolution's weights and bias.
* Every k=3, stride 1, pad 1 Conv1d / ConvTranspose1d becomes
      out[t] = bias + sum_{tap=0..2} in[t + tap - 1] @ W[tap]          (zeros outside the window)
  ConvTranspose1d weight w[i,o,k]:  W[tap][i][o] = w[i,o,2-tap]   
  Conv1d          weight w[o,i,k]:  W[tap][i][o] = w[o,i,tap]
  and its bwd-data is the same form with  Wb[tap][o][i] = W[2-tap][i][o].
* decoder_input (Linear 2048->5120) is followed by decoder.0's ConvTranspose1d with no
  activation in between (SeqConvVAE.py:134-136), so the two linear maps are composed into one
  [latent] -> [T][256] matrix: 2.35x fewer decoder FLOPs per closure evaluation, same function.
* torch.flatten orders the encoder output c*T+t (SeqConvVAE.py:105); activations here are
  token-major (t*512+c), so fc_mu / fc_var rows are permuted accordingly and concatenated
  (mu | logvar) into a single [T*512] -> [2*latent] layer.
"""
from __future__ import annotations

import numpy as np
import torch

BN_EPS = 1e-5


def _t64(v, device):
    if isinstance(v, torch.Tensor):
        return v.detach().to(device=device, dtype=torch.float64)
    return torch.from_numpy(np.array(v)).to(device=device, dtype=torch.float64)


def _bn_fold(sd, prefix, device):
    g, b = _t64(sd[prefix + ".weight"], device), _t64(sd[prefix + ".bias"], device)
    m, v = _t64(sd[prefix + ".running_mean"], device), _t64(sd[prefix + ".running_var"], device)
    scale = g / torch.sqrt(v + BN_EPS)
    return scale, b - m * scale


def _convT_taps(w):      # (in,out,3) -> [tap][in][out]
    return torch.stack([w[:, :, 2 - tap] for tap in range(3)], dim=0)


def _conv_taps(w):       # (out,in,3) -> [tap][in][out]
    return torch.stack([w[:, :, tap].transpose(0, 1) for tap in range(3)], dim=0)


def _bwd_taps(wf):       # [tap][in][out] -> [tap][out][in]
    return torch.stack([wf[2 - tap].transpose(0, 1) for tap in range(3)], dim=0)


def _pad_cols(w, mult=4):
    n = w.shape[-1]
    pad = (-n) % mult
    if pad:
        w = torch.nn.functional.pad(w, (0, pad))
    return w


class PreparedVae:
    """fp32 device tensors for one VAE + the layer table (taps, k, n) in execution order."""

    def __init__(self, state_dict, device, seq_len=10, latent_dim=2048, channels=45):
        dev = torch.device(device)
        T = seq_len
        sd = state_dict
        self.device = dev
        self.seq_len, self.latent_dim, self.channels = T, latent_dim, channels
        self.tensors = []          # keeps device memory alive
        self.dec, self.dec_bwd, self.enc = [], [], []

        def layer(w, bias, taps, k, n):
            w32 = _pad_cols(w).to(torch.float32).contiguous()
            assert w32.shape == (taps, k, (n + 3) // 4 * 4), (w32.shape, taps, k, n)
            b32 = None if bias is None else bias.to(torch.float32).contiguous()
            self.tensors += [w32, b32]
            return dict(w=w32, bias=b32, taps=taps, k=k, n=n)

        # ---------------- decoder, forward ----------------
        fwd = []      # folded [tap][in][out] float64 and bias for the five convs after the Linear
        for i in range(4):
            s, sh = _bn_fold(sd, f"decoder.{i}.1", dev)
            w = _convT_taps(_t64(sd[f"decoder.{i}.0.weight"], dev)) * s[None, None, :]
            fwd.append((w, _t64(sd[f"decoder.{i}.0.bias"], dev) * s + sh))
        s, sh = _bn_fold(sd, "final_layer.1", dev)
        w = _convT_taps(_t64(sd["final_layer.0.weight"], dev)) * s[None, None, :]
        fwd.append((w, _t64(sd["final_layer.0.bias"], dev) * s + sh))
        fwd.append((_conv_taps(_t64(sd["final_layer.3.weight"], dev)), _t64(sd["final_layer.3.bias"], dev)))

        # fuse decoder_input with decoder.0's ConvT: [latent] -> [T][256]
        Wl = _t64(sd["decoder_input.weight"], dev).view(512, T, latent_dim)      # [i][t'][n]
        bl = _t64(sd["decoder_input.bias"], dev).view(512, T)
        w1, b1 = fwd[0]                                                           # [tap][512][256]
        c1 = w1.shape[2]
        Wc = torch.zeros(latent_dim, T, c1, dtype=torch.float64, device=dev)
        bc = b1[None, :].repeat(T, 1)
        for t in range(T):
            for tap in range(3):
                tp = t + tap - 1
                if 0 <= tp < T:
                    Wc[:, t, :] += Wl[:, tp, :].transpose(0, 1) @ w1[tap]
                    bc[t] += bl[:, tp] @ w1[tap]
        Wc = Wc.reshape(1, latent_dim, T * c1)
        self.dec.append(layer(Wc, bc.reshape(-1), 1, latent_dim, T * c1))
        for w, b in fwd[1:]:
            self.dec.append(layer(w, b, 3, w.shape[1], w.shape[2]))

        # ---------------- decoder, bwd-data (execution order: last layer first) ----------------
        for w, _ in reversed(fwd[1:]):
            self.dec_bwd.append(layer(_bwd_taps(w), None, 3, w.shape[2], w.shape[1]))
        self.dec_bwd.append(layer(Wc[0].transpose(0, 1).reshape(1, T * c1, latent_dim), None, 1, T * c1, latent_dim))

        # ---------------- encoder ----------------
        for i in range(5):
            s, sh = _bn_fold(sd, f"encoder.{i}.1", dev)
            w = _conv_taps(_t64(sd[f"encoder.{i}.0.weight"], dev)) * s[None, None, :]
            self.enc.append(layer(w, _t64(sd[f"encoder.{i}.0.bias"], dev) * s + sh, 3, w.shape[1], w.shape[2]))
        cenc = 512
        wm = _t64(sd["fc_mu.weight"], dev).view(latent_dim, cenc, T)            # [n][c][t]
        wv = _t64(sd["fc_var.weight"], dev).view(latent_dim, cenc, T)
        wfc = torch.cat([wm, wv], dim=0).permute(2, 1, 0).reshape(1, T * cenc, 2 * latent_dim)   # [(t,c)][n]
        bfc = torch.cat([_t64(sd["fc_mu.bias"], dev), _t64(sd["fc_var.bias"], dev)])
        self.enc.append(layer(wfc, bfc, 1, T * cenc, 2 * latent_dim))

    def nbytes(self):
        return sum(t.numel() * 4 for t in self.tensors if t is not None)
