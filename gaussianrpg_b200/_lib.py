"""ctypes view of libgrpg_b200.so -- the C-ABI declared in include/grpg_b200.h.

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing
the product path raises.  (The CPU oracle under oracle/ is test infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libgrpg_b200.so"


class GeomLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "total_bytes", "rec", "depth_key", "rect", "tiles_touched", "cov3d", "clamped", "sorted_idx", "offsets",
        "scratch", "tile_mask", "num_rendered")]


class BinningLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("total_bytes", "point_list", "tile_keys", "scratch")]


class ImageLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("total_bytes", "n_contrib", "ranges")]


_fp = C.c_void_p  # device pointers cross the ABI as plain addresses


class ForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("S", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int), ("debug", C.c_int),
        ("background", _fp), ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("semantics", _fp),
        ("opacities", _fp), ("scales", _fp), ("rotations", _fp), ("cov3D_precomp", _fp),
        ("viewmatrix", _fp), ("projmatrix", _fp), ("cam_pos", _fp),
        ("out_color", _fp), ("out_depth", _fp), ("out_alpha", _fp), ("out_semantic", _fp), ("radii", _fp),
        ("geom_ws", _fp), ("binning_ws", _fp), ("image_ws", _fp), ("stream", _fp),
        ("tile_row_stride", C.c_int), ("tile_row_phase", C.c_int), ("forward_only", C.c_int), ("reference_binning", C.c_int),
        ("n_peer_frames", C.c_int), ("peer_frames", _fp * 8),
    ]


class BackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("S", C.c_int), ("R", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float), ("debug", C.c_int),
        ("background", _fp), ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("semantics", _fp),
        ("alphas", _fp), ("scales", _fp), ("rotations", _fp), ("cov3D_precomp", _fp),
        ("viewmatrix", _fp), ("projmatrix", _fp), ("cam_pos", _fp), ("radii", _fp),
        ("geom_ws", _fp), ("binning_ws", _fp), ("image_ws", _fp),
        ("dL_dpix", _fp), ("dL_dpix_depth", _fp), ("dL_dalphas", _fp), ("dL_dpix_semantic", _fp),
        ("dL_dmean2D", _fp), ("dL_dconic", _fp), ("dL_dopacity", _fp), ("dL_dcolor", _fp), ("dL_ddepth", _fp),
        ("dL_dmean3D", _fp), ("dL_dcov3D", _fp), ("dL_dsh", _fp), ("dL_dscale", _fp), ("dL_drot", _fp),
        ("dL_dsemantic", _fp), ("grad_ws", _fp), ("stream", _fp),
        ("tile_row_stride", C.c_int), ("tile_row_phase", C.c_int), ("stages", C.c_int), ("p_begin", C.c_int),
        ("p_count", C.c_int), ("n_peer_grad", C.c_int), ("peer_grad_ws", _fp * 8),
        ("pixel_grads_full_frame", C.c_int), ("reserved_", C.c_int),
    ]


class L1SsimArgs(C.Structure):  # grpg_l1_ssim_args (include/grpg_loss.h)
    _fields_ = [
        ("planes", C.c_int), ("planes_per_mask", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("img1", _fp), ("img2", _fp), ("mask", _fp), ("coef_l1", C.c_float), ("coef_ssim", C.c_float),
        ("sums", _fp), ("grad", _fp), ("stream", _fp),
    ]


class Rgb8Args(C.Structure):  # grpg_rgb8_args (include/grpg_image.h)
    _fields_ = [
        ("height", C.c_int), ("width", C.c_int), ("rgb", _fp), ("acc", _fp), ("sky", _fp), ("out_rgb8", _fp),
        ("out_rgb", _fp), ("stream", _fp),
    ]


COMPOSE_MAX_FOURIER = 16


class ComposeSubmodel(C.Structure):  # grpg_compose_submodel (include/grpg_compose.h)
    _fields_ = [
        ("n", C.c_int), ("is_actor", C.c_int), ("fourier_dim", C.c_int), ("reserved", C.c_int),
        ("xyz", _fp), ("scaling", _fp), ("rotation", _fp), ("opacity", _fp), ("features_dc", _fp),
        ("features_rest", _fp), ("flip_mask", _fp),
        ("obj_rot", _fp), ("obj_trans", _fp), ("idft", C.c_float * COMPOSE_MAX_FOURIER),
        ("d_xyz", _fp), ("d_scaling", _fp), ("d_rotation", _fp), ("d_opacity", _fp), ("d_features_dc", _fp),
        ("d_features_rest", _fp),
    ]


class AdamTensor(C.Structure):  # grpg_adam_tensor (include/grpg_optim.h)
    _fields_ = [
        ("param", _fp), ("grad", _fp), ("exp_avg", _fp), ("exp_avg_sq", _fp), ("numel", C.c_longlong),
        ("one_minus_beta1", C.c_float), ("beta2", C.c_float), ("one_minus_beta2", C.c_float), ("step_size", C.c_float),
        ("bias_correction2_sqrt", C.c_float), ("eps", C.c_float),
    ]


class StatsSubmodel(C.Structure):  # grpg_stats_submodel (include/grpg_optim.h)
    _fields_ = [("n", C.c_int), ("reserved", C.c_int), ("max_radii2D", _fp), ("xyz_gradient_accum", _fp), ("denom", _fp)]


class SkyArgs(C.Structure):  # grpg_sky_args (include/grpg_sky.h)
    _fields_ = [("height", C.c_int), ("width", C.c_int), ("resolution", C.c_int), ("cubemap", _fp), ("ray_matrix", _fp),
                ("jitter", _fp), ("mask", _fp), ("acc", _fp), ("fill", C.c_float), ("sky", _fp), ("stream", _fp)]


class ExchangeArgs(C.Structure):  # grpg_exchange_args (include/grpg_b200.h)
    _fields_ = [("P", C.c_int), ("world", C.c_int), ("rank", C.c_int), ("geom_ws", _fp), ("grad_rec", _fp),
                ("inbox", _fp * 8), ("stream", _fp)]


# every symbol include/*.h declare, with its ctypes signature
SYMBOLS = {
    "grpg_get_geometry_layout": (C.c_int, [C.c_int, C.POINTER(GeomLayout)]),
    "grpg_get_binning_layout": (C.c_int, [C.c_longlong, C.POINTER(BinningLayout)]),
    "grpg_get_image_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(ImageLayout)]),
    "grpg_band_rows": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "grpg_band_height": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "grpg_forward_geometry": (C.c_int, [C.POINTER(ForwardArgs), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "grpg_forward_render": (C.c_int, [C.POINTER(ForwardArgs), C.c_int]),
    "grpg_forward": (C.c_int, [C.POINTER(ForwardArgs), C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "grpg_forward_static": (C.c_int, [C.POINTER(ForwardArgs), C.c_longlong, _fp]),
    "grpg_exchange_inbox_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "grpg_exchange_pack": (C.c_int, [C.POINTER(ExchangeArgs)]),
    "grpg_exchange_accumulate": (C.c_int, [C.POINTER(ExchangeArgs)]),
    "grpg_backward_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "grpg_backward": (C.c_int, [C.POINTER(BackwardArgs)]),
    "grpg_mark_visible": (C.c_int, [C.c_int, _fp, _fp, _fp, _fp, _fp]),
    "grpg_visible_filter": (C.c_int, [C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_float, _fp, _fp, _fp, _fp,
                                      C.c_float, C.c_float, _fp, _fp, _fp]),
    "grpg_debug_reference_keys": (C.c_int, [C.c_int, C.c_longlong, _fp, _fp, _fp, _fp]),
    "grpg_debug_packed_math_check": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, _fp, _fp]),
    "grpg_profile_begin": (C.c_int, []),
    "grpg_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
    "grpg_last_error": (C.c_char_p, []),
    "grpg_version": (C.c_int, []),
    "grpg_l1_ssim": (C.c_int, [C.POINTER(L1SsimArgs)]),
    "grpg_compose_rgb8": (C.c_int, [C.POINTER(Rgb8Args)]),
    "grpg_sky_forward": (C.c_int, [C.POINTER(SkyArgs)]),
    "grpg_sky_backward": (C.c_int, [C.POINTER(SkyArgs), _fp, _fp]),
    "grpg_sky_compose_rgb8": (C.c_int, [C.POINTER(SkyArgs), _fp, _fp, _fp]),
    "grpg_knn_workspace_bytes": (C.c_size_t, [C.c_int]),
    "grpg_knn_mean_dist2": (C.c_int, [C.c_int, _fp, _fp, _fp, _fp]),
    "grpg_adam_workspace_bytes": (C.c_size_t, [C.c_int]),
    "grpg_adam_step": (C.c_int, [C.POINTER(AdamTensor), C.c_int, _fp, _fp]),
    "grpg_stats_workspace_bytes": (C.c_size_t, [C.c_int]),
    "grpg_densify_stats": (C.c_int, [C.POINTER(StatsSubmodel), C.c_int, _fp, _fp, _fp, _fp]),
    "grpg_densify_stats_ex": (C.c_int, [C.POINTER(StatsSubmodel), C.c_int, _fp, _fp, _fp, C.c_int, _fp, _fp]),
    "grpg_compose_workspace_bytes": (C.c_size_t, [C.c_int]),
    "grpg_compose_forward": (C.c_int, [C.POINTER(ComposeSubmodel), C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "grpg_compose_backward": (C.c_int, [C.POINTER(ComposeSubmodel), C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                        _fp, _fp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library once; build it in-tree first when sources are newer (needs nvcc)."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("GRPG_NO_AUTOBUILD", "0") != "1":
        try:
            from . import build as _build
            if _build.needs_build():
                _build.build()
        except Exception as exc:  # no nvcc on this box: fall through to the prebuilt .so
            if not LIB_PATH.exists():
                raise RuntimeError(f"libgrpg_b200.so is not built and cannot be built here: {exc}") from exc
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} missing -- run `python -m gaussianrpg_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().grpg_last_error().decode("utf-8", "replace")


def current_stream_ptr(dev) -> int:
    """Raw cudaStream_t of PyTorch's current stream on `dev` (the Stream-object route costs ~25 us per call)."""
    import torch
    idx = dev.index if getattr(dev, "index", None) is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)
