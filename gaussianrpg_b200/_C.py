"""Host-side mirror of the reference's pybind module `diff_gaussian_rasterization._C`.

Same four entry points, argument order and return tuples as
submodules/diff-gaussian-rasterization/ext.cpp:15-20 + rasterize_points.h:18-88, implemented on
top of the C-ABI (include/grpg_b200.h) through ctypes.  PyTorch is used only for device memory
and the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch

from . import _lib

NUM_CHANNELS = 3
_NEED_BINNING = 2   # GRPG_NEED_BINNING (include/grpg_b200.h)
_last_binned = {}    # (P, W, H, band, device) -> instances binned by the previous call: sizes the next workspace

# ---- static-capacity mode (grpg_forward_static): no host synchronisation, CUDA-graph capturable --------------------
_static_capacity = None   # instances; None = the default path (one host sync per forward, exact workspace)
_static_counts = {}       # key -> pinned int64[4] (num_binned, num_rendered, overflow, depth-sorted Gaussians) of the
#                           last static forward with that key; written asynchronously on the forward's stream


def set_static_binning(capacity):
    """Switch every following forward of this process to the sync-free path with room for `capacity` Gaussian/tile
    instances (None switches back).  The forward then never waits for the device -- the reference blocks on a
    device-to-host copy of num_rendered (rasterizer_impl.cu:284), and so does the default path here -- which makes
    a whole training step capturable into a CUDA graph.  A frame that needs more instances than `capacity` is rendered
    incompletely and flagged: call `check_static_binning()` after a synchronisation point (e.g. right after reading
    the loss) -- it raises with the capacity that was needed."""
    global _static_capacity
    _static_capacity = None if capacity is None else int(capacity)
    return _static_capacity


def static_binning_counts():
    """{key: (num_binned, num_rendered, overflow, depth_sorted)} of the last static forwards; valid after the stream
    they ran on has been synchronised."""
    return {k: tuple(int(x) for x in v) for k, v in _static_counts.items()}


def check_static_binning():
    """Raise if any static forward since the last call overflowed its capacity (call after a synchronisation)."""
    for k, v in _static_counts.items():
        if int(v[2]) != 0:
            raise RuntimeError(f"static binning capacity {_static_capacity} exceeded: the frame needs {int(v[0])} "
                               f"instances (forward key {k}); call set_static_binning() with a larger capacity")


_cuda_ok = None


def _require_cuda():
    global _cuda_ok
    if _cuda_ok is None:  # torch.cuda.is_available() costs ~10 us per call (NVML probe): ask once
        _cuda_ok = bool(torch.cuda.is_available())
    if not _cuda_ok:
        raise RuntimeError("gaussianrpg_b200 needs a CUDA device (sm_100a); there is no CPU path")


def _f32(t: torch.Tensor) -> torch.Tensor:
    """contiguous float32 view on the GPU, as the reference's `.contiguous().data<float>()` demands."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t) -> int | None:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _stream(dev=None) -> int:
    """Raw handle of the current PyTorch stream (torch.cuda.current_stream() builds a Stream object through
    several Python layers, ~25 us; the raw query is sub-microsecond)."""
    idx = dev.index if dev is not None and dev.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(_lib.last_error())


def _layouts(P: int, R: int, W: int, H: int):
    lib = _lib.load()
    g, b, i = _lib.GeomLayout(), _lib.BinningLayout(), _lib.ImageLayout()
    lib.grpg_get_geometry_layout(P, C.byref(g))
    lib.grpg_get_binning_layout(R, C.byref(b))
    lib.grpg_get_image_layout(W, H, C.byref(i))
    return g, b, i


def rasterize_gaussians(background, means3D, colors, semantics, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh,
                        degree, campos, prefiltered, debug, *, _band=(1, 0), _forward_only=False, _peer_frames=None,
                        _reference_binning=False, _static=None) -> Tuple[int, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                   torch.Tensor, torch.Tensor, torch.Tensor]:
    """RasterizeGaussiansCUDA (rasterize_points.cu:35-124).

    Returns (num_rendered, color[3,H,W], depth[1,H,W], alpha[1,H,W], semantic[S,H,W], radii[P],
    geomBuffer, binningBuffer, imgBuffer).

    `_band=(stride, phase)` (keyword-only, not part of the reference surface) restricts the call to the tile
    rows r with r % stride == phase; the images then use the compact band layout [C, rows*16, W]
    (gaussianrpg_b200.dist reassembles them).  `_reference_binning=True` bins exactly the reference's tile
    rectangles (index buffers then equal the reference's element for element); by default rectangles are clipped
    to the alpha >= 1/255 footprint, which drops instances the reference sorts and then skips.  `num_rendered` is
    the reference's count either way.  `_static=capacity` (default: the process-wide `set_static_binning` value) runs
    the sync-free path; the returned num_rendered is then `capacity` (an upper bound that grpg_backward accepts), the
    true counts land in `static_binning_counts()`.
    """
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:58-60
    _require_cuda()
    lib = _lib.load()
    dev = means3D.device
    P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
    S = int(semantics.shape[1]) if semantics is not None and semantics.ndim == 2 else 0
    M = int(sh.shape[1]) if sh is not None and sh.numel() != 0 else 0

    f32 = dict(dtype=torch.float32, device=dev)
    u8 = dict(dtype=torch.uint8, device=dev)
    if P == 0:  # rasterize_points.cu:70-74,86
        return (0, torch.zeros((NUM_CHANNELS, H, W), **f32), torch.zeros((1, H, W), **f32),
                torch.zeros((1, H, W), **f32), torch.zeros((S, H, W), **f32),
                torch.zeros((0,), dtype=torch.int32, device=dev), torch.empty(0, **u8), torch.empty(0, **u8),
                torch.empty(0, **u8))

    means3D_c, opacity_c = _f32(means3D), _f32(opacity)
    keep = [means3D_c, opacity_c]  # keeps converted temporaries alive until the launches are queued

    def opt(t):
        if t is None or t.numel() == 0:
            return None
        t = _f32(t)
        keep.append(t)
        return t

    sh_c, colors_c, sem_c = opt(sh), opt(colors), opt(semantics)
    scales_c, rot_c, cov_c = opt(scales), opt(rotations), opt(cov3D_precomp)
    bg_c, view_c, proj_c, cam_c = _f32(background), _f32(viewmatrix), _f32(projmatrix), _f32(campos)
    keep += [bg_c, view_c, proj_c, cam_c]

    stride, phase = int(_band[0]), int(_band[1])
    HL = lib.grpg_band_height(H, stride, phase)
    # every pixel / Gaussian row is written by the kernels: no zero fill needed (band images are padded to
    # whole tiles, so zero them when the last tile row is ragged)
    alloc = torch.zeros if (stride > 1 and H % 16 != 0) else torch.empty
    out_color = alloc((NUM_CHANNELS, HL, W), **f32)
    out_depth = alloc((1, HL, W), **f32)
    out_alpha = alloc((1, HL, W), **f32)
    out_semantic = alloc((S, HL, W), **f32)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)

    gl, _, il = _layouts(P, 0, W, HL)
    geom = torch.empty(gl.total_bytes, **u8)
    img = torch.empty(il.total_bytes, **u8)

    a = _lib.ForwardArgs()
    a.P, a.D, a.M, a.S = P, int(degree), M, S
    a.width, a.height = W, H
    a.tan_fovx, a.tan_fovy, a.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
    a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
    a.background, a.means3D, a.shs, a.colors_precomp = _ptr(bg_c), _ptr(means3D_c), _ptr(sh_c), _ptr(colors_c)
    a.semantics, a.opacities = _ptr(sem_c), _ptr(opacity_c)
    a.scales, a.rotations, a.cov3D_precomp = _ptr(scales_c), _ptr(rot_c), _ptr(cov_c)
    a.viewmatrix, a.projmatrix, a.cam_pos = _ptr(view_c), _ptr(proj_c), _ptr(cam_c)
    a.out_color, a.out_depth, a.out_alpha = out_color.data_ptr(), out_depth.data_ptr(), out_alpha.data_ptr()
    a.out_semantic = _ptr(out_semantic)
    a.radii = radii.data_ptr()
    a.geom_ws, a.image_ws, a.binning_ws = geom.data_ptr(), img.data_ptr(), None
    a.stream = _stream(dev)
    a.tile_row_stride, a.tile_row_phase = stride, phase
    a.forward_only = int(bool(_forward_only))
    a.reference_binning = int(bool(_reference_binning))
    if _peer_frames:
        a.n_peer_frames = len(_peer_frames)
        for i, ptr in enumerate(_peer_frames):
            a.peer_frames[i] = int(ptr)

    key = (P, W, H, stride, phase, bool(_reference_binning), dev.index)
    bl = _lib.BinningLayout()
    static_cap = _static_capacity if _static is None else (int(_static) or None)
    if static_cap and not debug:
        lib.grpg_get_binning_layout(static_cap, C.byref(bl))
        binning = torch.empty(max(int(bl.total_bytes), 256), **u8)
        a.binning_ws = _ptr(binning)
        counts = _static_counts.get(key)
        if counts is None:
            counts = _static_counts[key] = torch.zeros(4, dtype=torch.int64).pin_memory()
            if len(_static_counts) > 64:
                _static_counts.pop(next(iter(_static_counts)))
        with torch.cuda.device(dev):
            _check(lib.grpg_forward_static(C.byref(a), static_cap, counts.data_ptr()))
        return static_cap, out_color, out_depth, out_alpha, out_semantic, radii, geom, binning, img

    # The binning workspace is sized by a device-side count.  Guess it from the previous call with the same shape
    # (+25 %), so that both stages run inside one library call and the render stage is launched right behind the
    # host sync; when the guess is too small the library stops after the geometry stage and the exact size is used.
    guess = _last_binned.get(key, 0)
    lib.grpg_get_binning_layout(int(guess * 1.25) + 4096, C.byref(bl))
    binning = torch.empty(max(int(bl.total_bytes), 256), **u8)  # never empty: backward needs a valid pointer
    a.binning_ws = _ptr(binning)
    with torch.cuda.device(dev):
        n_binned, n_rendered = C.c_int(0), C.c_int(0)
        rc = lib.grpg_forward(C.byref(a), binning.numel(), C.byref(n_binned), C.byref(n_rendered))
        if rc == _NEED_BINNING:
            lib.grpg_get_binning_layout(int(n_binned.value), C.byref(bl))
            binning = torch.empty(max(int(bl.total_bytes), 256), **u8)
            a.binning_ws = _ptr(binning)
            rc = lib.grpg_forward_render(C.byref(a), int(n_binned.value))
        _check(rc)
    _last_binned[key] = int(n_binned.value)
    if len(_last_binned) > 64:
        _last_binned.pop(next(iter(_last_binned)))
    R = int(n_rendered.value)
    return R, out_color, out_depth, out_alpha, out_semantic, radii, geom, binning, img


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                 dL_dout_depth, dL_dout_alpha, dL_dout_semantic, sh, degree, campos, geomBuffer, R,
                                 binningBuffer, imageBuffer, alphas, semantics, debug, *, _band=(1, 0), _height=None,
                                 _width=None, _stage=3, _grad_rec=None, _slice=None, _peer_grad=None, _wanted=None,
                                 _full_frame_grads=False, _grad_rec_rows=None, _full_rows=False):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:126-220).

    Returns (dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dopacity[P,1], dL_dmeans3D[P,3], dL_dcov3D[P,6],
    dL_dsh[P,M,3], dL_dscales[P,3], dL_drotations[P,4], dL_dsemantic[P,S]).

    Keyword-only extras for the multi-GPU path (gaussianrpg_b200.dist): `_band` as in the forward (the pixel
    gradients and `alphas` are then band images and `_height`/`_width` are the frame size); `_stage=1` runs only the
    blend backward and returns (grad_rec[P,12], dL_dsemantic[P,S]); `_stage=2` runs only the per-Gaussian
    backward for `_slice=(p_begin, p_count)` from `_grad_rec[p_count,12]` and returns p_count-row gradients;
    with `_peer_grad` (peer-mapped pointers to every rank's full [P,12] record buffer) stage 2 sums the records
    over the ranks while loading them instead of reading `_grad_rec`.  `_wanted=(means2D, colors, cov3D, scale_rot)`
    (booleans; the autograd function passes `ctx.needs_input_grad`) skips allocating and writing the outputs nobody
    reads -- they come back as None -- as well as the two internal ones (dL_dconic, dL_ddepth).
    `_full_frame_grads=True` (with a band): the four pixel-gradient images are full frames [C,H,W], read by the kernel
    at the band's true rows, instead of compact band images.  `_grad_rec_rows` (stage 1): allocate the record buffer
    with that many rows (>= P, the surplus zeroed) so that it can be reduce-scattered without a padding copy.
    `_full_rows=True` (stage 2 on a slice): the gradients come back with all P rows, zero outside the slice.
    """
    _require_cuda()
    lib = _lib.load()
    dev = means3D.device
    P = int(means3D.shape[0])
    stride, phase = int(_band[0]), int(_band[1])
    if _stage & 1:
        H, W = int(dL_dout_color.shape[1]), int(dL_dout_color.shape[2])
        S = int(dL_dout_semantic.shape[0])
    else:
        H, W = 16, 16  # unused by the geometry stage
        S = int(semantics.shape[1]) if semantics is not None and semantics.ndim == 2 else 0
    if _height is not None:
        H = int(_height)
    if _width is not None:
        W = int(_width)
    M = int(sh.shape[1]) if sh is not None and sh.numel() != 0 else 0
    f32 = dict(dtype=torch.float32, device=dev)
    p_begin, p_count = (0, P) if _slice is None else (int(_slice[0]), int(_slice[1]))
    full_rows = bool(_full_rows) and _stage == 2 and (p_begin != 0 or p_count != P)
    Pout = P if (_stage != 2 or full_rows) else p_count
    row0 = p_begin if full_rows else 0  # the kernel writes rows [0, p_count) behind the pointers it is given

    def out(*shape):  # fully written by grpg_backward (a slice of a full-row result: zero elsewhere)
        return torch.zeros(shape, **f32) if (P == 0 or full_rows) else torch.empty(shape, **f32)

    def at(t):  # device pointer of row `row0`
        if t is None or t.numel() == 0:
            return None
        return t.data_ptr() + row0 * t.stride(0) * 4

    if _stage == 1:
        Pout = 0  # no per-parameter outputs in the blend-only stage
    P_full = P
    P = Pout
    w_m2d, w_col, w_cov, w_sr = (True, True, True, True) if _wanted is None else (bool(x) for x in _wanted)
    internal = _wanted is None  # the raw entry point mirrors the reference and materialises everything
    dL_dmeans3D, dL_dopacity, dL_dsh = out(P, 3), out(P, 1), out(P, M, 3)
    dL_dmeans2D = out(P, 3) if w_m2d else None
    dL_dcolors = out(P, NUM_CHANNELS) if w_col else None
    dL_ddepths, dL_dconic = (out(P, 1), out(P, 2, 2)) if internal else (None, None)
    dL_dcov3D = out(P, 6) if w_cov else None
    dL_dscales, dL_drot = (out(P, 3), out(P, 4)) if w_sr else (None, None)
    dL_dsemantic = out(P_full if _stage != 2 else P, S)
    P = P_full
    if P == 0:
        return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drot,
                dL_dsemantic)

    keep = []

    def opt(t):
        if t is None or t.numel() == 0:
            return None
        t = _f32(t)
        keep.append(t)
        return t

    a = _lib.BackwardArgs()
    a.P, a.D, a.M, a.S, a.R = P, int(degree), M, S, int(R)
    a.width, a.height = W, H
    a.tan_fovx, a.tan_fovy, a.scale_modifier, a.debug = float(tan_fovx), float(tan_fovy), float(scale_modifier), int(
        bool(debug))
    a.background, a.means3D = _ptr(opt(background)), _ptr(opt(means3D))
    a.shs, a.colors_precomp, a.semantics = _ptr(opt(sh)), _ptr(opt(colors)), _ptr(opt(semantics))
    a.alphas = _ptr(opt(alphas))
    a.scales, a.rotations, a.cov3D_precomp = _ptr(opt(scales)), _ptr(opt(rotations)), _ptr(opt(cov3D_precomp))
    a.viewmatrix, a.projmatrix, a.cam_pos = _ptr(opt(viewmatrix)), _ptr(opt(projmatrix)), _ptr(opt(campos))
    radii_c = radii.contiguous()
    a.radii = radii_c.data_ptr()
    a.geom_ws, a.binning_ws, a.image_ws = _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer)
    a.dL_dpix, a.dL_dpix_depth = _ptr(opt(dL_dout_color)), _ptr(opt(dL_dout_depth))
    a.dL_dalphas, a.dL_dpix_semantic = _ptr(opt(dL_dout_alpha)), _ptr(opt(dL_dout_semantic))
    a.dL_dmean2D, a.dL_dconic, a.dL_dopacity = at(dL_dmeans2D), at(dL_dconic), at(dL_dopacity)
    a.dL_dcolor, a.dL_ddepth, a.dL_dmean3D = at(dL_dcolors), at(dL_ddepths), at(dL_dmeans3D)
    a.dL_dcov3D, a.dL_dsh = at(dL_dcov3D), at(dL_dsh)
    a.dL_dscale, a.dL_drot, a.dL_dsemantic = at(dL_dscales), at(dL_drot), _ptr(dL_dsemantic)
    if _grad_rec is not None:
        grad_ws = _grad_rec.contiguous()
    else:
        rows = max(P, int(_grad_rec_rows or 0))
        grad_ws = torch.empty((rows, 12), dtype=torch.float32, device=dev)
        if rows > P:
            grad_ws[P:].zero_()
    a.grad_ws = grad_ws.data_ptr()
    a.stream = _stream(dev)
    a.tile_row_stride, a.tile_row_phase = stride, phase
    a.stages, a.p_begin, a.p_count = int(_stage), p_begin, p_count
    a.pixel_grads_full_frame = int(bool(_full_frame_grads))
    if _peer_grad:
        a.n_peer_grad = len(_peer_grad)
        for i, ptr in enumerate(_peer_grad):
            a.peer_grad_ws[i] = int(ptr)
    with torch.cuda.device(dev):
        _check(lib.grpg_backward(C.byref(a)))
    if _stage == 1:
        return grad_ws, dL_dsemantic
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drot, dL_dsemantic


def mark_visible(means3D, viewmatrix, projmatrix) -> torch.Tensor:
    """markVisible (rasterize_points.cu:222-242): bool[P], true where view-space z > 0.2."""
    _require_cuda()
    lib = _lib.load()
    P = int(means3D.shape[0])
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        m, v, p = _f32(means3D), _f32(viewmatrix), _f32(projmatrix)
        with torch.cuda.device(means3D.device):
            _check(lib.grpg_mark_visible(P, m.data_ptr(), v.data_ptr(), p.data_ptr(), present.data_ptr(), _stream()))
    return present


def rasterize_gaussians_filter(means3D, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                               tan_fovx, tan_fovy, image_height, image_width, prefiltered, debug):
    """RasterizeGaussiansfilterCUDA (rasterize_points.cu:244-306): (radii[P] int32, means2D[P,2])."""
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    _require_cuda()
    lib = _lib.load()
    dev = means3D.device
    P = int(means3D.shape[0])
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    means2D = torch.zeros((P, 2), dtype=torch.float32, device=dev)
    if P != 0:
        m, v, p = _f32(means3D), _f32(viewmatrix), _f32(projmatrix)
        s = _f32(scales) if scales is not None and scales.numel() else None
        r = _f32(rotations) if rotations is not None and rotations.numel() else None
        c = _f32(cov3D_precomp) if cov3D_precomp is not None and cov3D_precomp.numel() else None
        with torch.cuda.device(dev):
            _check(lib.grpg_visible_filter(P, int(image_width), int(image_height), m.data_ptr(), _ptr(s),
                                           float(scale_modifier), _ptr(r), _ptr(c), v.data_ptr(), p.data_ptr(),
                                           float(tan_fovx), float(tan_fovy), radii.data_ptr(), means2D.data_ptr(),
                                           _stream()))
    return radii, means2D
