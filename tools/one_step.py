"""A few forward+backward steps of the 2 M-Gaussian scene through the raw op (profiling target for ncu).

    python tools/one_step.py [steps] [P] [stride]

stride > 1: one interleaved tile-row band (phase 0) of a stride-way sharded frame instead of the whole frame -- forward and
the blend stage of the backward, what rank 0 of `bench.py --gpus stride` launches (the band kernels: split layout, batched)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import cases
from gaussianrpg_b200 import synthetic, _C
from gaussianrpg_b200 import dist as gd

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
sc_cpu = synthetic.street_scene(P=P)
sc = sc_cpu.to(dev)
dL = [t.to(dev) for t in cases.loss_grads(sc_cpu)]
if k == 1:
    for _ in range(steps):
        fwd = cases.raw_forward(_C, sc)
        g = cases.raw_backward(_C, sc, fwd, dL)
else:
    E = torch.Tensor([])
    sem = torch.zeros(P, 0, device=dev)
    gb = gd.frame_to_band(torch.cat([dL[0], dL[1], dL[2]], 0), k, 0)
    H, W = sc.height, sc.width
    for _ in range(steps):
        fwd = _C.rasterize_gaussians(sc.bg, sc.means3D, E, sem, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix,
                                     sc.projmatrix, sc.tanfovx, sc.tanfovy, H, W, sc.shs, sc.sh_degree, sc.campos, False,
                                     False, _band=(k, 0))
        g = _C.rasterize_gaussians_backward(sc.bg, sc.means3D, fwd[5], E, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix,
                                            sc.projmatrix, sc.tanfovx, sc.tanfovy, gb[:3].contiguous(), gb[3:4].contiguous(),
                                            gb[4:5].contiguous(), dL[3], sc.shs, sc.sh_degree, sc.campos, fwd[6], fwd[0],
                                            fwd[7], fwd[8], fwd[3], sem, False, _band=(k, 0), _height=H, _width=W, _stage=1)
torch.cuda.synchronize()
print("R", fwd[0])
