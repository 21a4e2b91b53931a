"""Host-side cost of one call through the public module on a scene whose GPU work is negligible
(P = 512, 64x64): what a per-step synchronising caller pays on top of the kernels."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from gaussianrpg_b200 import synthetic
import diff_gaussian_rasterization as dgr

dev = torch.device("cuda:0")
sc = synthetic.plumbing_scene(P=512, W=64, H=64, S=0, sh_degree=1, seed=1).to(dev)
leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
rast = dgr.GaussianRasterizer(sc.settings())

def fwd():
    with torch.no_grad():
        return rast(means3D=leaves["means3D"], means2D=None, opacities=leaves["opacities"], shs=leaves["shs"],
                    scales=leaves["scales"], rotations=leaves["rotations"])

def fwd_bwd():
    m2d = torch.zeros(512, 3, device=dev, requires_grad=True)
    out = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
               scales=leaves["scales"], rotations=leaves["rotations"])
    (out[0].mean() + out[2].mean() + out[3].mean()).backward()

for name, fn in (("forward", fwd), ("forward+backward", fwd_bwd)):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per call (host + tiny kernels)")
if len(sys.argv) > 1:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200):
        fwd_bwd()
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
