"""PyTorch operator surface of the rasterizer: settings tuple, autograd function, nn.Module.

Keeps the reference's public contract (submodules/diff-gaussian-rasterization/
diff_gaussian_rasterization/__init__.py): `GaussianRasterizationSettings` field names/order
(:167-179), `GaussianRasterizer.forward` keyword names, defaults and validation messages
(:197-233), the 5-tuple result `(color, radii, depth, alpha, semantic)` (:102), the gradient
tuple order (:152-163), `markVisible` (:186-195), `visible_filter` (:235-260), and the debug
snapshot files written when a call fails under `debug=True` (:87-94, :141-148).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _C


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _to_cpu(values):
    return tuple(v.detach().cpu().clone() if isinstance(v, torch.Tensor) else v for v in values)


def _call_with_snapshot(fn, args, debug: bool, dump_name: str, what: str):
    """Run fn(*args); under debug, keep a CPU copy of the arguments and dump it if the call throws."""
    if not debug:
        return fn(*args)
    saved = _to_cpu(args)
    try:
        return fn(*args)
    except Exception:
        torch.save(saved, dump_name)
        print(f"\nAn error occured in {what}. Please forward {dump_name} for debugging.")
        raise


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, grad_mode=True):
        rs = raster_settings
        native_args = (rs.bg, means3D, colors_precomp, semantics, opacities, scales, rotations, rs.scale_modifier,
                       cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                       rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        # When nothing can ask for a backward the backward-only state (cov3D, clamp flags) is not materialised.
        # `needs_input_grad` mirrors the tensors' requires_grad flags whatever the grad mode, and grad mode is always
        # off inside Function.forward, so the caller samples torch.is_grad_enabled() and passes it in (`grad_mode`).
        fwd_only = (not grad_mode) or not any(ctx.needs_input_grad)
        (num_rendered, color, depth, alpha, semantic, radii, geom, binning, img) = _call_with_snapshot(
            lambda *a_: _C.rasterize_gaussians(*a_, _forward_only=fwd_only), native_args, rs.debug, "snapshot_fw.dump",
            "forward")
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.tensor_inputs = [isinstance(t, torch.Tensor) for t in (means3D, means2D, sh, colors_precomp, semantics,
                                                                   opacities, scales, rotations, cov3Ds_precomp)]
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning,
                              img, alpha, semantics)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha, semantic

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha, grad_semantic):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img, alpha,
         semantics) = ctx.saved_tensors
        native_args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                       rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_color, grad_depth, grad_alpha,
                       grad_semantic, sh, rs.sh_degree, rs.campos, geom, ctx.num_rendered, binning, img, alpha,
                       semantics, rs.debug)
        # inputs of forward(): means3D 0, means2D 1, sh 2, colors_precomp 3, semantics 4, opacities 5, scales 6,
        # rotations 7, cov3Ds_precomp 8 -- gradients nobody needs are neither allocated nor written
        need = ctx.needs_input_grad
        wanted = (need[1], need[3], need[8], need[6] or need[7])
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot, g_sem) = _call_with_snapshot(
            lambda *a_: _C.rasterize_gaussians_backward(*a_, _wanted=wanted), native_args, rs.debug, "snapshot_bw.dump",
            "backward")
        # order of forward()'s inputs: means3D, means2D, sh, colors_precomp, semantics, opacities, scales,
        # rotations, cov3Ds_precomp, raster_settings
        grads = (g_means3D, g_means2D, g_sh, g_colors, g_sem, g_opac, g_scales, g_rot, g_cov3D)
        # autograd rejects a gradient for an argument that was not a tensor (e.g. means2D=None in eval)
        return tuple(g if is_t else None for g, is_t in zip(grads, ctx.tensor_inputs)) + (None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, torch.is_grad_enabled())


def _empty() -> torch.Tensor:
    return torch.Tensor([])


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    @torch.no_grad()
    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        rs = self.raster_settings
        return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs: Optional[torch.Tensor] = None,
                colors_precomp: Optional[torch.Tensor] = None, scales: Optional[torch.Tensor] = None,
                rotations: Optional[torch.Tensor] = None, cov3D_precomp: Optional[torch.Tensor] = None,
                semantics: Optional[torch.Tensor] = None):
        have_sh, have_col = shs is not None, colors_precomp is not None
        if have_sh == have_col:
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        have_sr = scales is not None and rotations is not None
        any_sr = scales is not None or rotations is not None
        have_cov = cov3D_precomp is not None
        if (not have_sr and not have_cov) or (any_sr and have_cov):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        shs = _empty() if shs is None else shs
        colors_precomp = _empty() if colors_precomp is None else colors_precomp
        scales = _empty() if scales is None else scales
        rotations = _empty() if rotations is None else rotations
        cov3D_precomp = _empty() if cov3D_precomp is None else cov3D_precomp
        if semantics is None:
            semantics = torch.zeros((means3D.shape[0], 0), dtype=torch.float32, device=means3D.device)
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, semantics, opacities, scales, rotations,
                                   cov3D_precomp, self.raster_settings)

    @torch.no_grad()
    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        scales = _empty() if scales is None else scales
        rotations = _empty() if rotations is None else rotations
        cov3D_precomp = _empty() if cov3D_precomp is None else cov3D_precomp
        return _C.rasterize_gaussians_filter(means3D, scales, rotations, rs.scale_modifier, cov3D_precomp,
                                             rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                                             rs.image_width, rs.prefiltered, rs.debug)
