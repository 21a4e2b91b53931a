"""Tiny forward+backward through the public surface (used under compute-sanitizer)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaussianrpg_b200 import synthetic
from diff_gaussian_rasterization import GaussianRasterizer

S = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = synthetic.plumbing_scene(P=512, W=80, H=60, S=S).to("cuda")
kw = sc.raster_kwargs()
for k in ("means3D", "opacities", "shs", "scales", "rotations", "semantics"):
    if kw[k] is not None:
        kw[k] = kw[k].clone().requires_grad_(True)
means2D = torch.zeros(sc.means3D.shape[0], 3, device="cuda", requires_grad=True)
rast = GaussianRasterizer(sc.settings(debug=True))
color, radii, depth, alpha, sem = rast(means2D=means2D, **kw)
loss = color.sum() + 0.1 * depth.sum() + 0.3 * alpha.sum() + (sem.sum() if sem.numel() else 0)
loss.backward()
torch.cuda.synchronize()
print("ok", float(loss), int((radii > 0).sum()), float(means2D.grad.abs().max()), float(kw["means3D"].grad.abs().max()))
