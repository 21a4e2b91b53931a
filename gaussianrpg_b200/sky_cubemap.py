"""Sky cube-map lookup: host-side mirror of `SkyCubeMap.forward`
(/root/reference/lib/models/sky_cubemap.py:77-124) over the C-ABI of include/grpg_sky.h.

    sky = sky_color(sky_cube_map, H, W, K, R, T, acc=acc)            # [3,H,W], clamped, autograd -> sky_cube_map
    rgb8, rgb = sky_compose_rgb8(sky_cube_map, H, W, K, R, T, rgb, acc)   # lookup + composite + clamp + uint8 HWC

`K` [3,3] intrinsics, `R`, `T` the world-to-camera rotation / translation the reference takes from
`camera.world_view_transform.transpose(0, 1)` (:91-92).  The reference evaluates the lookup with nvdiffrast
(`dr.texture(..., filter_mode='linear', boundary_mode='cube')`), which is not vendored in its tree; the cube mapping
restated here is documented in include/grpg_sky.h (parity unpinned).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def _ray_matrix(K: torch.Tensor, R: torch.Tensor) -> torch.Tensor:
    """rays_d of get_rays_torch (graphics_utils.py:186-207) is, before normalisation, R^T K^-1 (x, y, 1): the camera
    centre cancels.  Computed on the device (no host read); float32 like the reference's matmuls."""
    return (R.float().t() @ torch.inverse(K.float())).contiguous().reshape(9)


def _args(cube, H, W, M, jitter, mask, acc, fill, sky, dev):
    a = _lib.SkyArgs()
    a.height, a.width, a.resolution = int(H), int(W), int(cube.shape[1])
    a.cubemap, a.ray_matrix = cube.data_ptr(), M.data_ptr()
    a.jitter = jitter.data_ptr() if jitter is not None else None
    a.mask = mask.data_ptr() if mask is not None else None
    a.acc = acc.data_ptr() if acc is not None else None
    a.fill = float(fill)
    a.sky = sky.data_ptr() if sky is not None else None
    a.stream = _lib.current_stream_ptr(dev)
    return a


def _prep(cube, H, W, acc, mask, jitter):
    if not cube.is_cuda:
        raise RuntimeError("gaussianrpg_b200.sky_cubemap has no CPU path: the cube map must be a CUDA tensor")
    if cube.ndim != 4 or cube.shape[0] != 6 or cube.shape[1] != cube.shape[2] or cube.shape[3] != 3:
        raise RuntimeError("sky_cube_map must have shape (6, res, res, 3)")
    cube_c = cube.detach().contiguous().float()
    acc_c = None if acc is None else acc.detach().contiguous().float()
    if acc_c is not None and acc_c.numel() != H * W:
        raise RuntimeError("acc must have H*W elements")
    mask_c = None
    if mask is not None:
        if tuple(mask.shape) != (H, W):
            raise RuntimeError("mask must be (H, W)")
        mask_c = (mask if mask.dtype in (torch.bool, torch.uint8) else mask != 0).contiguous()
    jit_c = None
    if jitter is not None:
        if tuple(jitter.shape) != (2, H, W):
            raise RuntimeError("jitter must be (2, H, W)")
        jit_c = jitter.detach().contiguous().float()
    return cube_c, acc_c, mask_c, jit_c


class _SkyLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cube, H, W, M, acc, mask, jitter, fill):
        cube_c, acc_c, mask_c, jit_c = _prep(cube, H, W, acc, mask, jitter)
        dev = cube.device
        sky = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        a = _args(cube_c, H, W, M, jit_c, mask_c, acc_c, fill, sky, dev)
        with torch.cuda.device(dev):
            if _lib.load().grpg_sky_forward(C.byref(a)) != 0:
                raise RuntimeError(_lib.last_error())
        ctx.save_for_backward(cube, M, acc_c if acc_c is not None else torch.empty(0), mask_c if mask_c is not None else torch.empty(0),
                              jit_c if jit_c is not None else torch.empty(0), sky)
        ctx.dims, ctx.fill = (H, W), fill
        return sky

    @staticmethod
    def backward(ctx, g):
        cube, M, acc_c, mask_c, jit_c, sky = ctx.saved_tensors
        H, W = ctx.dims
        dev = cube.device
        cube_c = cube.detach().contiguous().float()
        d_cube = torch.zeros_like(cube_c)
        a = _args(cube_c, H, W, M, jit_c if jit_c.numel() else None, mask_c if mask_c.numel() else None,
                  acc_c if acc_c.numel() else None, ctx.fill, sky, dev)
        with torch.cuda.device(dev):
            if _lib.load().grpg_sky_backward(C.byref(a), g.contiguous().float().data_ptr(), d_cube.data_ptr()) != 0:
                raise RuntimeError(_lib.last_error())
        return d_cube, None, None, None, None, None, None, None


def sky_color(sky_cube_map: torch.Tensor, H: int, W: int, K: torch.Tensor, R: torch.Tensor, T: Optional[torch.Tensor] = None,
              acc: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, white_background: bool = False,
              jitter: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`SkyCubeMap.forward` (sky_cubemap.py:77-124): [3,H,W] sky colour, clamped to [0,1].  `mask` [H,W] = the
    training-time sky mask (:80-82); without it `acc` [1,H,W] gives mask = (1 - acc) > 1e-3 (:83-84); with neither the
    sky is looked up everywhere (:97-100).  Outside the mask the colour is 0 (1 with `white_background`, :104-115).
    `jitter` [2,H,W] = (perturb_i, perturb_j) of training mode (graphics_utils.py:194-197): pass
    `torch.rand(2, H, W)`; None = pixel centres.  `T` is accepted for signature parity; it cancels in the ray direction."""
    M = _ray_matrix(K.to(sky_cube_map.device), R.to(sky_cube_map.device))
    return _SkyLookup.apply(sky_cube_map, int(H), int(W), M, acc, mask, jitter, 1.0 if white_background else 0.0)


def sky_compose_rgb8(sky_cube_map: torch.Tensor, H: int, W: int, K: torch.Tensor, R: torch.Tensor, rgb: torch.Tensor,
                     acc: torch.Tensor, white_background: bool = False, out: Optional[torch.Tensor] = None,
                     return_float: bool = False):
    """What `StreetGaussianRenderer.render` does after the rasterizer in evaluation mode and what the simulator node
    does with the result, in ONE kernel: sky lookup (mask from acc) -> rgb + sky * (1 - acc) -> clamp -> x255 -> uint8
    HWC (street_gaussian_renderer.py:336-346, simulator.py:313-314).  Returns (bytes [H,W,3], float image or None)."""
    cube_c, acc_c, _, _ = _prep(sky_cube_map, H, W, acc, None, None)
    dev = sky_cube_map.device
    rgb_c = rgb.detach().contiguous().float()
    if tuple(rgb_c.shape) != (3, H, W):
        raise RuntimeError("rgb must have shape (3, H, W)")
    if out is None:
        out = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    elif tuple(out.shape) != (H, W, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
        raise RuntimeError("out must be a contiguous uint8 tensor of shape (H, W, 3)")
    elif not out.is_cuda and not out.is_pinned():
        raise RuntimeError("a host `out` must be pinned memory")
    outf = torch.empty_like(rgb_c) if return_float else None
    M = _ray_matrix(K.to(dev), R.to(dev))
    a = _args(cube_c, H, W, M, None, None, acc_c, 1.0 if white_background else 0.0, None, dev)
    with torch.cuda.device(dev):
        if _lib.load().grpg_sky_compose_rgb8(C.byref(a), rgb_c.data_ptr(), out.data_ptr(),
                                             outf.data_ptr() if outf is not None else None) != 0:
            raise RuntimeError(_lib.last_error())
    return out, outf
