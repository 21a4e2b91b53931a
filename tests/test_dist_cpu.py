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


# ---- the sharded operator's host logic with a CPU stand-in for the native module ------------------------------------
class _MockNative:
    """`_C`-shaped stand-in (pure torch, CPU): a 'render' and a staged 'backward' that are linear in their inputs, so
    that the orchestration of _ShardedRasterize -- bands, frame assembly, record exchange, slices, zero fill -- can be
    checked against a single-process evaluation under gloo."""

    @staticmethod
    def _frame(means3D, H, W):
        yy = torch.arange(H, dtype=torch.float32)[:, None]
        xx = torch.arange(W, dtype=torch.float32)[None, :]
        s = means3D.sum()
        return torch.stack([s * (yy + 1) + xx, s * yy - xx, s + yy * xx, s * 0.5 + yy + 0 * xx, s * 0.25 + xx + 0 * yy], 0)

    def rasterize_gaussians(self, bg, means3D, colors, semantics, opacity, scales, rotations, sm, cov, view, proj, tx, ty, H,
                            W, sh, deg, campos, pre, debug, *, _band=(1, 0), _peer_frames=None, _forward_only=False):
        k, r = _band
        f = gd.frame_to_band(self._frame(means3D.detach(), H, W), k, r)
        P = means3D.shape[0]
        u8 = torch.zeros(8, dtype=torch.uint8)
        return (7, f[:3].contiguous(), f[3:4].contiguous(), f[4:5].contiguous(), torch.zeros(0, f.shape[1], W),
                torch.arange(P, dtype=torch.int32), u8, u8, u8)

    def rasterize_gaussians_backward(self, bg, means3D, radii, colors, scales, rotations, sm, cov, view, proj, tx, ty, g_color,
                                     g_depth, g_alpha, g_sem, sh, deg, campos, geom, R, binning, img, alphas, semantics,
                                     debug, *, _band=(1, 0), _height=None, _width=None, _stage=3, _grad_rec=None, _slice=None,
                                     _wanted=None, _full_frame_grads=False, _grad_rec_rows=None, _full_rows=False):
        P = means3D.shape[0]
        k, r = _band
        if _stage == 1:
            g = torch.cat([g_color, g_depth, g_alpha], 0)
            if _full_frame_grads:
                g = gd.frame_to_band(g, k, r)
            rows = max(P, _grad_rec_rows or 0)
            rec = torch.zeros(rows, 12)
            base = torch.stack([g[c].sum() for c in range(5)])
            rec[:P, :5] = base[None, :] * (torch.arange(P, dtype=torch.float32)[:, None] + 1)
            return rec, torch.zeros(P, 0)
        b, c = _slice
        rec = _grad_rec[:c]
        rows, off = (P, b) if _full_rows else (c, 0)

        def out(w, src):
            t = torch.zeros(rows, *w)
            t[off:off + c] = src.reshape(c, *w)
            return t
        M = sh.shape[1]
        return (out((3,), rec[:, :3]), None, out((1,), rec[:, 3:4]), out((3,), rec[:, 2:5]), None,
                out((M, 3), rec[:, :1].repeat(1, M * 3)), out((3,), rec[:, 1:4] * 2), out((4,), rec[:, :4] * 3),
                torch.zeros(rows, 0))


def _sharded_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gaussianrpg_b200.rasterizer import GaussianRasterizationSettings
        gd._native = _MockNative()
        P, H, W = 11, 70, 24  # ragged: 5 tile rows over 2 ranks, P not a multiple of the world size
        g = torch.Generator().manual_seed(3)
        base = dict(means3D=torch.randn(P, 3, generator=g), opacities=torch.rand(P, 1, generator=g),
                    shs=torch.randn(P, 4, 3, generator=g), scales=torch.rand(P, 3, generator=g),
                    rotations=torch.randn(P, 4, generator=g))
        rs = GaussianRasterizationSettings(H, W, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 1,
                                           torch.zeros(3), False, False)
        wgt = torch.randn(5, H, W, generator=g)

        def run(output, gradients, k_world):
            leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
            if k_world == 1:  # single-process evaluation of the same mock: the expected frame and gradients
                m = _MockNative()
                frame = m._frame(leaves["means3D"].detach(), H, W)
                rec, _ = m.rasterize_gaussians_backward(None, leaves["means3D"], None, None, None, None, 1, None, None, None, 1,
                                                        1, wgt[:3], wgt[3:4], wgt[4:5], None, leaves["shs"], 1, None, None, 7,
                                                        None, None, None, None, False, _stage=1)
                full = m.rasterize_gaussians_backward(None, leaves["means3D"], None, None, None, None, 1, None, None, None, 1,
                                                      1, None, None, None, None, leaves["shs"], 1, None, None, 7, None, None,
                                                      None, None, False, _stage=2, _grad_rec=rec, _slice=(0, P))
                return frame, dict(means3D=full[3], opacities=full[2], shs=full[5], scales=full[6], rotations=full[7])
            rast = gd.ShardedGaussianRasterizer(rs, output=output, gradients=gradients)
            color, radii, depth, alpha, sem = rast(means2D=None, **leaves)
            w = gd.frame_to_band(wgt, world, rank) if output == "band" else wgt
            ((color * w[:3]).sum() + (depth * w[3:4]).sum() + (alpha * w[4:5]).sum()).backward()
            return torch.cat([color, depth, alpha], 0).detach(), {k: v.grad for k, v in leaves.items()}

        frame_want, grads_want = run(None, None, 1)
        for output in ("frame", "band"):
            for gradients in ("full", "shard"):
                got, grads = run(output, gradients, world)
                want = frame_want if output == "frame" else gd.frame_to_band(frame_want, world, rank)
                assert torch.equal(got, want), (output, gradients)
                b, c = gd.gaussian_slice(P, world, rank)
                for k in grads_want:
                    if gradients == "full":
                        assert torch.allclose(grads[k], grads_want[k], rtol=1e-5, atol=1e-5), (output, gradients, k)
                    else:  # complete on the rank's slice, zero elsewhere; the sum over ranks is the full gradient
                        assert torch.allclose(grads[k][b:b + c], grads_want[k][b:b + c], rtol=1e-5, atol=1e-5), (output, k)
                        assert not grads[k][:b].any() and not grads[k][b + c:].any()
                        tot = grads[k].clone()
                        dist.all_reduce(tot)
                        assert torch.allclose(tot, grads_want[k], rtol=1e-5, atol=1e-5)
    finally:
        gd._native = None
        dist.destroy_process_group()


def test_sharded_operator_host_logic_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_sharded_worker, args=(2, port), nprocs=2, join=True)
