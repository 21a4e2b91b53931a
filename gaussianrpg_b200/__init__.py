"""gaussianrpg_b200 -- B200-native (sm_100a) differentiable 3D-Gaussian rasterizer.

Layout: csrc/ (CUDA kernels + the C-ABI of include/grpg_b200.h), _lib.py (ctypes loader),
_C.py (mirror of the reference's pybind module), rasterizer.py (PyTorch operator surface),
debug.py (test hooks into the opaque workspaces), dist.py (multi-GPU tile-row sharding).
"""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401
from ._C import set_static_binning, check_static_binning, static_binning_counts  # noqa: F401

__version__ = "0.1.0"
