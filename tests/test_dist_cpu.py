"""Host-side logic of the multi-GPU path on CPU: pure band/slice helpers, and the collectives under gloo
with world_size 2 (the kernels themselves are covered by the single-GPU band-emulation test)."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussianrpg_b200 import dist as gd


@pytest.mark.parametrize("H,k", [(1280, 8), (1066, 2), (1066, 4), (47, 3), (16, 4)])
def test_band_roundtrip_single_process(H, k):
    g = torch.Generator().manual_seed(H * 31 + k)
    x = torch.randn(5, H, 33, generator=g)
    bands = [gd.pad_band(gd.frame_to_band(x, k, r), H, k) for r in range(k)]
    assert sum(gd.band_rows(H, k, r) for r in range(k)) == gd.tile_rows(H)
    y = gd.bands_to_frame(torch.stack(bands), H)
    assert torch.equal(x, y)
    # a band holds exactly the tile rows t % k == r
    r = k - 1
    b = gd.frame_to_band(x, k, r)
    for i in range(gd.band_rows(H, k, r)):
        t = i * k + r
        rows = x[:, t * 16:(t + 1) * 16]
        assert torch.equal(b[:, i * 16:i * 16 + rows.shape[1]], rows)


def test_gaussian_slices_cover_everything():
    for P, k in [(10, 4), (2_000_000, 8), (7, 8), (0, 2)]:
        got = []
        for r in range(k):
            b, c = gd.gaussian_slice(P, k, r)
            got += list(range(b, b + c))
        assert got == list(range(P))
        assert gd.padded_count(P, k) % k == 0 and gd.padded_count(P, k) >= P


def _worker(rank, world, port, H):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        frame = torch.randn(5, H, 40, generator=g)               # same on every rank
        mine = gd.pad_band(gd.frame_to_band(frame, world, rank), H, world)
        gathered = gd._all_gather_rows(mine[None], dist.group.WORLD, world)
        assert torch.equal(gd.bands_to_frame(gathered, H), frame)
        # reduce-scatter of per-Gaussian records + all-gather of the shards == plain sum
        P = 37
        part = torch.randn(P, 12, generator=torch.Generator().manual_seed(100 + rank))
        total = sum(torch.randn(P, 12, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
        Pp = gd.padded_count(P, world)
        padded = torch.nn.functional.pad(part, (0, 0, 0, Pp - P))
        shard = gd._reduce_scatter_rows(padded, dist.group.WORLD, world, rank)
        b, c = gd.gaussian_slice(P, world, rank)
        assert torch.allclose(shard[:c], total[b:b + c], atol=1e-6)
        full = gd._all_gather_rows(shard, dist.group.WORLD, world)[:P]
        assert torch.allclose(full, total, atol=1e-6)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H", [64, 1066])
def test_collectives_gloo_world2(H):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, H), nprocs=2, join=True)
