"""`distCUDA2` of the reference's simple-knn submodule (/root/reference/submodules/simple-knn/spatial.cu:15-27) over the
C-ABI of include/grpg_knn.h: mean squared distance of every point to its three nearest other points.  The package
`simple_knn` at the repository root re-exports it as `simple_knn._C.distCUDA2`, the name the reference imports
(lib/models/gaussian_model.py:5,63).  There is no CPU path."""
from __future__ import annotations

import torch

from . import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """points [P,3] (CUDA) -> float32 [P]."""
    if not points.is_cuda:
        raise RuntimeError("gaussianrpg_b200.simple_knn has no CPU path: points must be a CUDA tensor")
    if points.ndim != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    P = int(points.shape[0])
    pts = points.detach().contiguous().float()
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)  # spatial.cu:21
    if P == 0:
        return means
    lib = _lib.load()
    ws = torch.empty(int(lib.grpg_knn_workspace_bytes(P)), dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        if lib.grpg_knn_mean_dist2(P, pts.data_ptr(), means.data_ptr(), ws.data_ptr(),
                                   _lib.current_stream_ptr(points.device)) != 0:
            raise RuntimeError(_lib.last_error())
    return means
