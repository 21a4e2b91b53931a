"""Drop-in `diff_gaussian_rasterization` package backed by the B200-native rasterizer.

Callers of the reference (lib/utils/camera_utils.py:13, lib/models/gaussian_renderer.py:3,
script/test_gaussian_rasterization.py:4) import exactly these names; the implementation lives in
`gaussianrpg_b200` (hand-written sm_100a CUDA behind a C-ABI).  `_C` mirrors the reference's
pybind module so code that reaches for `diff_gaussian_rasterization._C.*` keeps working.
"""
from gaussianrpg_b200 import _C  # noqa: F401
from gaussianrpg_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    rasterize_gaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
