"""Multi-GPU sharding of the rasterizer over one node (one process per GPU, torch.distributed / NCCL).

The reference is single-GPU (no comms anywhere); this is new work following SURVEY 8(e) and
BASELINE.json's north_star: the path shards where it does so naturally --

  forward   screen-tile rows.  Rank r of k owns the tile rows t with t % k == r (interleaved, because street
            scenes pile instances around the horizon and contiguous bands would be badly balanced).  Every
            rank projects all P Gaussians (cheaper than shipping 48 B records), but emits, sorts and blends
            only the instances of its rows; the k compact band images are all-gathered (NCCL) and interleaved
            back into the frame.
  backward  per-Gaussian gradient shards.  Each rank replays its own band and accumulates a partial 48-byte
            2D-gradient record per Gaussian; one reduce-scatter (sum) over the Gaussian axis leaves rank r
            with the complete records of Gaussians [r*P/k, (r+1)*P/k), on which it runs the per-Gaussian
            backward; the parameter-gradient shards are all-gathered so every rank returns full gradients
            (replicated optimiser, like the reference's single process).

Everything below the collectives goes through the same C-ABI as the single-GPU path
(`tile_row_stride/phase`, `stages`, `p_begin/p_count` in include/grpg_b200.h).
The pure tensor helpers (band <-> frame, slices) are device-agnostic and are what the gloo CPU tests cover.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist
from torch import nn

TILE = 16


# ---- pure helpers (CPU-testable) ------------------------------------------------------------------
def tile_rows(H: int) -> int:
    return (H + TILE - 1) // TILE


def band_rows(H: int, k: int, r: int) -> int:
    rows = tile_rows(H)
    if k <= 1:
        return rows
    return (rows - 1 - r) // k + 1 if r < rows else 0


def max_band_rows(H: int, k: int) -> int:
    return band_rows(H, k, 0)


def frame_to_band(full: torch.Tensor, k: int, r: int) -> torch.Tensor:
    """[C,H,W] frame -> compact band image [C, band_rows*16, W] of rank r (zero padded below the frame)."""
    C, H, W = full.shape
    lr_max = max_band_rows(H, k)
    pad = lr_max * k * TILE - H
    x = torch.nn.functional.pad(full, (0, 0, 0, pad)) if pad else full
    x = x.reshape(C, lr_max, k, TILE, W)[:, :, r]
    return x.reshape(C, lr_max * TILE, W)[:, : band_rows(H, k, r) * TILE].contiguous()


def pad_band(band: torch.Tensor, H: int, k: int) -> torch.Tensor:
    """pad a rank's band image to the common height max_band_rows*16 so it can be all-gathered"""
    want = max_band_rows(H, k) * TILE
    have = band.shape[1]
    return band if have == want else torch.nn.functional.pad(band, (0, 0, 0, want - have))


def bands_to_frame(gathered: torch.Tensor, H: int) -> torch.Tensor:
    """[k, C, max_band_rows*16, W] (rank-major all-gather result) -> [C,H,W] frame"""
    k, C, HLm, W = gathered.shape
    lr_max = HLm // TILE
    x = gathered.reshape(k, C, lr_max, TILE, W).permute(1, 2, 0, 3, 4).reshape(C, lr_max * k * TILE, W)
    return x[:, :H].contiguous()


def gaussian_slice(P: int, k: int, r: int) -> Tuple[int, int]:
    """equal-size slices of the padded Gaussian axis; the last ones may be short or empty"""
    per = (P + k - 1) // k
    b = min(P, r * per)
    return b, max(0, min(P, b + per) - b)


def padded_count(P: int, k: int) -> int:
    return ((P + k - 1) // k) * k


def _reduce_scatter_rows(x: torch.Tensor, group, k: int, r: int) -> torch.Tensor:
    """sum over ranks, keep this rank's row block of a tensor whose row count is a multiple of k"""
    per = x.shape[0] // k
    if dist.get_backend(group) == "nccl":
        out = torch.empty((per,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, x.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)  # gloo fallback used by the CPU tests
    return y[r * per:(r + 1) * per].contiguous()


def _all_gather_rows(x: torch.Tensor, group, k: int) -> torch.Tensor:
    out = torch.empty((k * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    else:
        parts = [torch.empty_like(x) for _ in range(k)]
        dist.all_gather(parts, x.contiguous(), group=group)
        out = torch.cat(parts, 0)
    return out


_EXCHANGE_PLAN = "replicate"  # or "shard"; see _ShardedRasterize.backward


class _PeerFrame:
    """Peer-mapped full-frame output buffer (torch symmetric memory over NVLink).  With it the blend kernel of every
    rank stores its band's pixels directly into all ranks' frames -- the forward all-gather becomes the kernel's
    epilogue (peer stores) instead of a separate NCCL collective plus an interleave copy."""
    _cache = {}

    def __init__(self, channels: int, H: int, W: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.frame = symm_mem.empty((channels, H, W), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.frame, group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]

    @classmethod
    def get(cls, channels, H, W, device, group):
        key = (channels, H, W, str(device), id(group))
        if key not in cls._cache:
            try:
                cls._cache[key] = cls(channels, H, W, device, group)
            except Exception as exc:  # no peer mapping on this system: use the NCCL all-gather
                cls._cache[key] = None
                if dist.get_rank(group) == 0:
                    print(f"[gaussianrpg_b200.dist] symmetric memory unavailable ({exc!r}); using NCCL all-gather")
        return cls._cache[key]


class _PeerRecords:
    """Peer-mapped [P,12] gradient-record buffer: every rank's blend backward accumulates into its own copy, and
    the per-Gaussian backward of every rank sums all copies in its load path (no separate all-reduce)."""
    _cache = {}

    def __init__(self, P: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.rec = symm_mem.empty((P, 12), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.rec, group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]

    @classmethod
    def get(cls, P, device, group):
        key = (P, str(device), id(group))
        if key not in cls._cache:
            try:
                cls._cache[key] = cls(P, device, group)
            except Exception as exc:
                cls._cache[key] = None
                if dist.get_rank(group) == 0:
                    print(f"[gaussianrpg_b200.dist] symmetric memory unavailable ({exc!r}); using NCCL all-reduce")
        return cls._cache[key]


FUSED_FORWARD_GATHER = True  # set False to force the NCCL all-gather path
# Sum the gradient records inside the per-Gaussian backward (peer loads) for groups up to this size.  Reading
# k-1 peers directly costs (k-1) x 48 B per visible Gaussian against 2(k-1)/k x 48 B per Gaussian for the ring
# all-reduce, so the direct form only wins (fewer bytes, one kernel less, no extra pass over HBM) for small k.
FUSED_RECORD_REDUCE_MAX_RANKS = 0


# ---- the sharded operator --------------------------------------------------------------------------
class _ShardedRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, group):
        from . import _C
        rs = raster_settings
        k, r = dist.get_world_size(group), dist.get_rank(group)
        H, W = rs.image_height, rs.image_width
        S_in = int(semantics.shape[1]) if semantics is not None and semantics.ndim == 2 else 0
        peer = None
        if FUSED_FORWARD_GATHER and k > 1 and k <= 8 and means3D.is_cuda and dist.get_backend(group) == "nccl":
            peer = _PeerFrame.get(5 + S_in, H, W, means3D.device, group)
        if peer is not None:
            peer.handle.barrier(channel=0)  # every rank has finished reading the previous frame
        out = _C.rasterize_gaussians(rs.bg, means3D, colors_precomp, semantics, opacities, scales, rotations,
                                     rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                                     rs.tanfovy, H, W, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug,
                                     _band=(k, r), _peer_frames=peer.ptrs if peer is not None else None)
        R, color_b, depth_b, alpha_b, sem_b, radii, geom, binning, img = out
        S = sem_b.shape[0]
        if peer is not None:
            peer.handle.barrier(channel=1)  # every rank's band has landed in this rank's frame
            frame = peer.frame.clone()
        else:
            packed = pad_band(torch.cat([color_b, depth_b, alpha_b, sem_b], 0), H, k)  # one collective for all planes
            gathered = _all_gather_rows(packed[None], group, k)
            frame = bands_to_frame(gathered, H)
        color, depth, alpha, sem = frame[:3], frame[3:4], frame[4:5], frame[5:5 + S]
        ctx.rs, ctx.group, ctx.R = rs, group, R
        ctx.tensor_inputs = [isinstance(t, torch.Tensor) for t in (means3D, means2D, sh, colors_precomp, semantics,
                                                                   opacities, scales, rotations, cov3Ds_precomp)]
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning,
                              img, alpha_b, semantics)
        ctx.mark_non_differentiable(radii)
        return color.contiguous(), radii, depth.contiguous(), alpha.contiguous(), sem.contiguous()

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha, g_sem):
        from . import _C
        rs, group = ctx.rs, ctx.group
        k, r = dist.get_world_size(group), dist.get_rank(group)
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img, alpha_b,
         semantics) = ctx.saved_tensors
        H, W = rs.image_height, rs.image_width
        P = means3D.shape[0]
        S = g_sem.shape[0]
        # the blend backward reads the band's rows straight out of the full-frame loss gradients
        common = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                  rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy)
        tail = (sh, rs.sh_degree, rs.campos, geom, ctx.R, binning, img, alpha_b, semantics, rs.debug)
        peer_rec = None
        if (_EXCHANGE_PLAN == "replicate" and 1 < k <= FUSED_RECORD_REDUCE_MAX_RANKS and means3D.is_cuda
                and dist.get_backend(group) == "nccl"):
            peer_rec = _PeerRecords.get(P, means3D.device, group)
        grad_rec, g_semantics = _C.rasterize_gaussians_backward(
            *common, g_color.contiguous(), g_depth.contiguous(), g_alpha.contiguous(), g_sem.contiguous(), *tail,
            _band=(k, r), _height=H, _width=W, _stage=1, _full_frame_grads=True,
            _grad_rec=peer_rec.rec if peer_rec is not None else None)
        # Sum the per-Gaussian 2D gradient records over ranks.  Two exchange plans:
        #  "shard":     reduce-scatter the 48 B records, per-Gaussian backward on the owned slice, all-gather the
        #               parameter-gradient shards (104 B per Gaussian for means/sh/opacity/scale/rotation);
        #  "replicate": reduce-scatter + all-gather of the 48 B records themselves (= all-reduce), then every rank
        #               runs the (cheap, 0.15 ms at 2 M) per-Gaussian backward for all Gaussians.
        # "replicate" moves less than half the bytes and is the default; "shard" is what a ZeRO-style sharded
        # optimiser would use (it would simply skip the final all-gather).
        if _EXCHANGE_PLAN == "replicate":
            if peer_rec is not None:
                peer_rec.handle.barrier(channel=2)  # every rank's records are complete
            else:
                dist.all_reduce(grad_rec, op=dist.ReduceOp.SUM, group=group)
            shard = _C.rasterize_gaussians_backward(
                *common, g_color, g_depth, g_alpha, g_sem, *tail, _band=(k, r), _height=H, _width=W, _stage=2,
                _grad_rec=grad_rec, _slice=(0, P), _peer_grad=peer_rec.ptrs if peer_rec is not None else None)
            if peer_rec is not None:
                peer_rec.handle.barrier(channel=3)  # nobody still reads this rank's records: they may be reused
            g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot = shard[:8]
        else:
            Pp = padded_count(P, k)
            if Pp != P:
                grad_rec = torch.nn.functional.pad(grad_rec, (0, 0, 0, Pp - P))
            mine = _reduce_scatter_rows(grad_rec, group, k, r)
            p_begin, p_count = gaussian_slice(P, k, r)
            shard = _C.rasterize_gaussians_backward(
                *common, g_color, g_depth, g_alpha, g_sem, *tail, _band=(k, r), _height=H, _width=W, _stage=2,
                _grad_rec=mine[:max(p_count, 1)], _slice=(p_begin, p_count))
            per = Pp // k
            full = []
            for t in shard[:8]:  # means2D, colors, opacity, means3D, cov3D, sh, scales, rot
                flat = t.reshape(t.shape[0], -1)
                if flat.shape[0] != per:
                    flat = torch.nn.functional.pad(flat, (0, 0, 0, per - flat.shape[0]))
                g = _all_gather_rows(flat, group, k)[:P]
                full.append(g.reshape((P,) + tuple(t.shape[1:])))
            g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot = full
        if S > 0:
            dist.all_reduce(g_semantics, op=dist.ReduceOp.SUM, group=group)
        grads = (g_means3D, g_means2D, g_sh, g_colors, g_semantics, g_opac, g_scales, g_rot, g_cov3D)
        return tuple(g if is_t else None for g, is_t in zip(grads, ctx.tensor_inputs)) + (None, None)


class ShardedGaussianRasterizer(nn.Module):
    """Drop-in for `GaussianRasterizer` when every rank of `group` holds the same Gaussians and camera:
    same arguments, same 5-tuple, same (full) gradients on every rank."""

    def __init__(self, raster_settings, group=None):
        super().__init__()
        self.raster_settings = raster_settings
        self.group = group if group is not None else dist.group.WORLD

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, semantics=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        E = torch.Tensor([])
        if semantics is None:
            semantics = torch.zeros((means3D.shape[0], 0), dtype=torch.float32, device=means3D.device)
        return _ShardedRasterize.apply(means3D, means2D, E if shs is None else shs,
                                       E if colors_precomp is None else colors_precomp, semantics, opacities,
                                       E if scales is None else scales, E if rotations is None else rotations,
                                       E if cov3D_precomp is None else cov3D_precomp, self.raster_settings, self.group)


# ---- bench.py --gpus N ------------------------------------------------------------------------------
def bench_sharded(args, sc_cpu, dev, rank: int, world: int) -> dict:
    """One frame of the 2 M-Gaussian scene split over `world` GPUs (strong scaling): forward + backward with the
    collectives inside the timed region, device-timed, max over ranks."""
    sc = sc_cpu.to(dev)
    H, W = sc.height, sc.width
    g = torch.Generator().manual_seed(123)
    gt = torch.rand(3, H, W, generator=g).to(dev)
    w_depth = (torch.rand(1, H, W, generator=g) * 0.01).to(dev)
    w_alpha = (torch.rand(1, H, W, generator=g) * 0.1).to(dev)
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    rast = ShardedGaussianRasterizer(sc.settings())
    P = sc.means3D.shape[0]

    def fwd_bwd():
        for v in leaves.values():
            v.grad = None
        means2D = torch.zeros(P, 3, device=dev, requires_grad=True)
        color, radii, depth, alpha, _ = rast(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                                             shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        loss = (color - gt).abs().mean() + (depth * w_depth).mean() + (alpha * w_alpha).mean()
        loss.backward()
        return loss

    def fwd_only():
        with torch.no_grad():
            return rast(means3D=leaves["means3D"], means2D=None, opacities=leaves["opacities"], shs=leaves["shs"],
                        scales=leaves["scales"], rotations=leaves["rotations"])

    def timed(fn, steps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(3, args.warmup)):
        fwd_bwd()
    ms_fb = timed(fwd_bwd, args.steps)
    for _ in range(3):
        fwd_only()
    ms_f = timed(fwd_only, args.steps)

    # end to end: per-step host inputs (camera + ground truth, pinned) in, loss scalar out
    gt_host = gt.cpu().pin_memory()

    def e2e():
        gt_d = gt_host.to(dev, non_blocking=True)
        for v in leaves.values():
            v.grad = None
        means2D = torch.zeros(P, 3, device=dev, requires_grad=True)
        color, radii, depth, alpha, _ = rast(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                                             shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        loss = (color - gt_d).abs().mean() + (depth * w_depth).mean() + (alpha * w_alpha).mean()
        loss.backward()
        return float(loss.item())

    for _ in range(2):
        e2e()
    ms_e2e = timed(e2e, args.steps)
    # camera-parallel companion line ("pure camera-parallel ... should be reported beside it"): every rank
    # renders its OWN camera of a batch of `world` cameras with the single-GPU operator, no exchange at all
    from .rasterizer import GaussianRasterizer
    solo = GaussianRasterizer(sc.settings())

    def solo_fwd_bwd():
        for v in leaves.values():
            v.grad = None
        means2D = torch.zeros(P, 3, device=dev, requires_grad=True)
        color, radii, depth, alpha, _ = solo(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                                             shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        ((color - gt).abs().mean() + (depth * w_depth).mean() + (alpha * w_alpha).mean()).backward()

    def solo_fwd():
        with torch.no_grad():
            solo(means3D=leaves["means3D"], means2D=None, opacities=leaves["opacities"], shs=leaves["shs"],
                 scales=leaves["scales"], rotations=leaves["rotations"])

    for _ in range(3):
        solo_fwd_bwd(); solo_fwd()
    ms_solo_fb = timed(solo_fwd_bwd, args.steps)
    ms_solo_f = timed(solo_fwd, args.steps)
    band_bytes = 5 * max_band_rows(H, world) * TILE * W * 4
    return {
        "metric": "iters_per_sec_fwd_bwd_1920x1280_2M", "value": 1000.0 * args.steps / ms_fb, "unit": "iters/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_fb / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "ours", "fwd_fps": 1000.0 * args.steps / ms_f,
        "e2e": {"value": 1000.0 * args.steps / ms_e2e, "unit": "iters/s", "h2d_bytes_per_step": int(gt_host.numel() * 4),
                "d2h_bytes_per_step": 4},
        # kernels of this library per step and rank: 22 forward (preprocess, 12 depth-sort, scan, emit, 5 tile-sort,
        # ranges, blend) + 2 backward -- the count bench.py measures with the library's profiler on one GPU
        "gpu_launches": 24 * args.steps * world,
        "camera_parallel": {"value": 1000.0 * args.steps * world / ms_solo_fb, "unit": "iters/s",
                            "fwd_fps": 1000.0 * args.steps * world / ms_solo_f, "scaling": "weak",
                            "note": f"a batch of {world} cameras, one per GPU, single-GPU operator, no exchange "
                                    f"(the upper bound SURVEY 8e asks to report beside the sharded number)"},
        "config": {"workload": f"street scene {P} Gaussians, {W}x{H}, fwd+bwd, one frame split over {world} GPUs",
                   "parallelism": f"tile rows interleaved mod {world} (fwd, all-gather {band_bytes} B/rank) + "
                                  f"all-reduce (reduce-scatter + all-gather) of the 48 B per-Gaussian gradient records "
                                  f"(bwd, {48 * P} B), per-Gaussian backward replicated",
                   "l2": "no flush: one step streams > 126 MB per GPU"},
    }
