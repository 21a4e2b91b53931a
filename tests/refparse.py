"""Typed views into the REFERENCE extension's opaque buffers (test helper).

Layout restated from cuda_rasterizer/rasterizer_impl.cu:155-193 (`fromChunk`) and
rasterizer_impl.h:40-60 (`obtain`: every sub-array starts at the next multiple of 128 bytes).
"""
from __future__ import annotations

import torch


def _carve(buf: torch.Tensor, fields):
    base = buf.data_ptr()
    off = 0
    out = {}
    for name, count, dtype in fields:
        addr = (base + off + 127) // 128 * 128
        off = addr - base
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        out[name] = buf[off:off + nbytes].view(dtype)
        off += nbytes
    return out


def parse_reference(P: int, R: int, W: int, H: int, geom: torch.Tensor, binning: torch.Tensor, img: torch.Tensor):
    g = _carve(geom, [("depths", P, torch.float32), ("clamped", 3 * P, torch.uint8), ("internal_radii", P, torch.int32),
                      ("means2D", 2 * P, torch.float32), ("cov3D", 6 * P, torch.float32),
                      ("conic_opacity", 4 * P, torch.float32), ("rgb", 3 * P, torch.float32),
                      ("tiles_touched", P, torch.int32)])
    out = dict(depths=g["depths"], clamped=g["clamped"].view(P, 3), means2D=g["means2D"].view(P, 2),
               cov3D=g["cov3D"].view(P, 6), conic_opacity=g["conic_opacity"].view(P, 4), rgb=g["rgb"].view(P, 3),
               tiles_touched=g["tiles_touched"])
    N = W * H
    i = _carve(img, [("n_contrib", N, torch.int32), ("ranges", 2 * N, torch.int32)])
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out["n_contrib"] = i["n_contrib"].view(H, W)
    out["ranges"] = i["ranges"].view(N, 2)[:tiles]
    if R > 0:
        b = _carve(binning, [("point_list", R, torch.int32), ("point_list_unsorted", R, torch.int32),
                             ("point_list_keys", R, torch.int64)])
        out["point_list"] = b["point_list"]
        out["point_list_keys"] = b["point_list_keys"]
    return out
