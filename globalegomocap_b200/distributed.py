"""Multi-GPU: one process per GPU, the windows of each sequence sharded in contiguous
blocks across ranks (reference optimizer.py:370 loops over them sequentially; they are
independent).  The only exchange is at stitching time: a rank needs its left neighbour's last
window and its right neighbour's first window (results, 10 x 45 float64 each) to average the two
overlap frames at the shard boundary (optimizer.py:425-437) and to give the sigma=1 Gaussian
(optimizer.py:450, radius 4) true neighbours.  That is one all-gather of two windows per rank and
sequence — a few KB over NVLink — and nothing inside the L-BFGS loop.

The arithmetic (merge, smoothing) is injected: the product passes Engine methods (CUDA), the
world_size-2 gloo tests on CPU pass the oracle's numpy functions.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world ranks get one more."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owned_frames(lo: int, hi: int, n_windows: int, stride: int, overlap: int):
    """Merged frames a shard owns: [stride*lo, stride*hi), plus the final `overlap` frames on the
    last shard, so that the shards tile the (stride*W + overlap)-frame sequence exactly."""
    return stride * lo, stride * hi + (overlap if hi == n_windows else 0)


def exchange_boundary_windows(windows, group=None):
    """windows: this rank's results [S, Wr, T, J, 3] float64 for S sequences (Wr >= 1), or a list of S tensors
    [Wr_s, T, J, 3] when the sequences differ in length.
    Returns (left, right): [S, T, J, 3] tensors holding the left neighbour's last window and the
    right neighbour's first window (None at the ends of the rank line)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return None, None
    if isinstance(windows, torch.Tensor):
        mine = torch.stack([windows[:, 0], windows[:, -1]], dim=1).contiguous()    # [S,2,T,J,3]
    else:
        mine = torch.stack([torch.stack([w[0], w[-1]]) for w in windows]).contiguous()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    left = gathered[rank - 1][:, 1] if rank > 0 else None
    right = gathered[rank + 1][:, 0] if rank < world - 1 else None
    return left, right


def stitch_shard(windows, left, right, merge_fn, smooth_fn, stride: int, overlap: int, final_smooth=True):
    """Merged (+ smoothed) frames owned by this shard of ONE sequence.

    windows [Wr,T,J,3]; left / right: neighbour windows [T,J,3] or None.  The neighbours' windows are
    appended, the extended run is merged and smoothed, and the owned slice is cut out: owned frames
    are >= stride frames away from the cut ends, the Gaussian's radius (4) is smaller, and the frames
    it reaches into (window interiors and the shared overlap) are final, so the result equals the
    single-process one bit for bit."""
    parts = ([left[None]] if left is not None else []) + [windows] + ([right[None]] if right is not None else [])
    ext = torch.cat(parts, dim=0) if len(parts) > 1 else windows
    seq = merge_fn(ext, overlap)
    if final_smooth:
        seq = smooth_fn(seq)
    start = stride if left is not None else 0
    n_own = stride * windows.shape[0] + (overlap if right is None else 0)
    return seq[start:start + n_own]


def stitch_optimized(engine, windows_per_sequence, final_smooth=True, group=None):
    """The optimised global sequence of every clip from this rank's windows (reference optimizer.py:421-450: merge
    with overlap averaging, then the sigma = 1 Gaussian), one all-gather of boundary windows when the clips' windows
    are sharded over ranks.  windows_per_sequence: list of float64 CUDA tensors [Wr_s, T, J, 3] in the global frame.
    Returns the list of this rank's owned frames per sequence (the whole sequence in a single process)."""
    stride, overlap = engine.T - 2, 2
    sharded = dist.is_initialized() and dist.get_world_size(group) > 1
    left, right = exchange_boundary_windows(windows_per_sequence, group) if sharded else (None, None)
    outs = []
    for s, wins in enumerate(windows_per_sequence):
        outs.append(stitch_shard(wins, None if left is None else left[s], None if right is None else right[s],
                                 lambda w, ov: engine.merge_windows(w, ov), lambda q: engine.gaussian_smooth(q, 1.0),
                                 stride, overlap, final_smooth=final_smooth))
    return outs
