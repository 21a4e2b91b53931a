"""Test hooks into the opaque workspaces returned by `_C.rasterize_gaussians`.

The internal layout is private to the library (include/grpg_b200.h, grpg_*_layout); this module
only turns the byte blobs into typed views so parity tests can compare tile/key indices
bit-exactly with the reference's GeometryState/BinningState/ImageState
(cuda_rasterizer/rasterizer_impl.cu:155-193).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _view(buf: torch.Tensor, offset: int, count: int, dtype: torch.dtype) -> torch.Tensor:
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    return buf[offset:offset + nbytes].view(dtype)


def parse_buffers(P: int, R: int, W: int, H: int, geom: torch.Tensor, binning: torch.Tensor, img: torch.Tensor):
    """`R` (the num_rendered returned by the forward) is only a fallback: the number of binned instances, which
    sizes the binning buffer, is read from the geometry buffer (it is smaller than num_rendered unless the
    forward ran with `_reference_binning=True`)."""
    lib = _lib.load()
    gl, bl, il = _lib.GeomLayout(), _lib.BinningLayout(), _lib.ImageLayout()
    lib.grpg_get_geometry_layout(P, C.byref(gl))
    if P > 0:
        counts = _view(geom, gl.num_rendered, 4, torch.int64).cpu()
        num_rendered, R, n_sorted = int(counts[1]), int(counts[0]), int(counts[3])
    else:
        num_rendered, n_sorted = R, 0
    lib.grpg_get_binning_layout(R, C.byref(bl))
    lib.grpg_get_image_layout(W, H, C.byref(il))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    rec = _view(geom, gl.rec, P * 12, torch.float32).view(P, 3, 4)
    rect = _view(geom, gl.rect, P * 2, torch.int32).view(P, 2)
    out = dict(
        means2D=rec[:, 0, 0:2], conic_opacity=rec[:, 1, :], rgb=rec[:, 2, 0:3],
        depths=rec[:, 2, 3], tiles_touched=_view(geom, gl.tiles_touched, P, torch.int32),
        cov3D=_view(geom, gl.cov3d, P * 6, torch.float32).view(P, 6), clamped=_view(geom, gl.clamped, P, torch.uint8),
        # depth order of the Gaussians that emit instances (tiles_touched != 0): the first n_sorted entries
        sorted_idx=_view(geom, gl.sorted_idx, P, torch.int32)[:n_sorted],
        offsets=_view(geom, gl.offsets, P, torch.int32)[:n_sorted],
        tile_mask=_view(geom, gl.tile_mask, P, torch.int32),
        rect_min=torch.stack([rect[:, 0] & 0xFFFF, rect[:, 1] & 0xFFFF], 1),
        rect_max=torch.stack([(rect[:, 0] >> 16) & 0xFFFF, (rect[:, 1] >> 16) & 0xFFFF], 1),
        n_contrib=_view(img, il.n_contrib, W * H, torch.int32).view(H, W),
        ranges=_view(img, il.ranges, tiles * 2, torch.int32).view(tiles, 2),
        num_binned=R, num_rendered=num_rendered,
    )
    if R > 0:
        out["point_list"] = _view(binning, bl.point_list, R, torch.int32)
        out["tile_keys"] = _view(binning, bl.tile_keys, R, torch.int32)
        keys = torch.empty(R, dtype=torch.int64, device=geom.device)
        rc = lib.grpg_debug_reference_keys(P, R, geom.data_ptr(), binning.data_ptr(), keys.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        out["point_list_keys"] = keys  # (tile << 32) | depth bits, as rasterizer_impl.cu:102-104
    else:
        out["point_list"] = torch.empty(0, dtype=torch.int32, device=geom.device)
        out["tile_keys"] = torch.empty(0, dtype=torch.int32, device=geom.device)
        out["point_list_keys"] = torch.empty(0, dtype=torch.int64, device=geom.device)
    return out
