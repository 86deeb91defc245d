"""Window sharding + boundary exchange for stitching, world_size 2 on CPU (gloo).  The arithmetic
(merge, smoothing) is injected: here the oracle's numpy functions, in the product the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from globalegomocap_b200 import distributed as gd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _merge(w, overlap):
    from oracle import pipeline_np as pl
    return torch.from_numpy(pl.merge_batches(w.numpy(), overlap))


def _smooth(q):
    from oracle import pipeline_np as pl
    return torch.from_numpy(pl.gaussian_filter1d_reflect(q.numpy(), 1.0, 0))


def _worker(rank, world, port, n_windows, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        S = 3
        full = torch.from_numpy(rng.standard_normal((S, n_windows, 10, 15, 3)))        # every rank draws the same data
        lo, hi = gd.shard_range(n_windows, rank, world)
        mine = full[:, lo:hi].contiguous()
        left, right = gd.exchange_boundary_windows(mine)
        assert (left is None) == (rank == 0) and (right is None) == (rank == world - 1)
        outs = [gd.stitch_shard(mine[s], None if left is None else left[s], None if right is None else right[s],
                                _merge, _smooth, 8, 2, final_smooth=True) for s in range(S)]
        f0, f1 = gd.owned_frames(lo, hi, n_windows, 8, 2)
        assert all(o.shape[0] == f1 - f0 for o in outs)
        torch.save((f0, f1, torch.stack(outs)), os.path.join(tmpdir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [2, 7, 12])
def test_sharded_stitching_equals_single_process(tmp_path, n_windows):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_windows, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    full = torch.from_numpy(rng.standard_normal((3, n_windows, 10, 15, 3)))
    ref = torch.stack([_smooth(_merge(full[s], 2)) for s in range(3)])
    parts = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == 8 * n_windows + 2
    got = torch.cat([p[2] for p in parts], dim=1)
    assert got.shape == ref.shape
    assert torch.equal(got, ref)            # bit-identical: same arithmetic on the same neighbours


def test_shard_ranges_tile_the_windows():
    for n in (1, 2, 7, 374, 1870):
        for world in (1, 2, 4, 8):
            r = [gd.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
            f = [gd.owned_frames(a, b, n, 8, 2) for a, b in r if b > a]
            assert f[0][0] == 0 and f[-1][1] == 8 * n + 2
