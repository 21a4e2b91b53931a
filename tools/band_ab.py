"""Per-kernel times of the forward and the blend backward on the bench scene for the whole frame (stride 1) and for one
interleaved tile-row band of a k-way sharded frame (stride k, phase 0: what rank 0 of `bench.py --gpus k` runs), on ONE
GPU.  Used to A/B blend-kernel variants (GRPG_FWD_PIPE / GRPG_BWD_PIPE, ...) for the band case without a multi-GPU
box.  Prints one JSON line per stride.

    python tools/band_ab.py [strides, default 1,8] [steps, default 10]
"""
import ctypes as C
import json
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import cases
from gaussianrpg_b200 import synthetic, _C, _lib
from gaussianrpg_b200 import dist as gd

strides = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,8").split(",")]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
sc_cpu = synthetic.street_scene()
sc = sc_cpu.to(dev)
H, W, P = sc.height, sc.width, sc.means3D.shape[0]
dL = [t.to(dev) for t in cases.loss_grads(sc_cpu)]
gcat = torch.cat([dL[0], dL[1], dL[2]], 0)
E = torch.Tensor([])
sem = torch.zeros(P, 0, device=dev)
lib = _lib.load()
env = {k: v for k, v in os.environ.items() if k.startswith("GRPG_")}
for k in strides:
    gb = gd.frame_to_band(gcat, k, 0) if k > 1 else gcat

    def step():
        o = _C.rasterize_gaussians(sc.bg, sc.means3D, E, sem, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix,
                                   sc.projmatrix, sc.tanfovx, sc.tanfovy, H, W, sc.shs, sc.sh_degree, sc.campos, False,
                                   False, _band=(k, 0))
        rec, _ = _C.rasterize_gaussians_backward(sc.bg, sc.means3D, o[5], E, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix,
                                                 sc.projmatrix, sc.tanfovx, sc.tanfovy, gb[:3].contiguous(),
                                                 gb[3:4].contiguous(), gb[4:5].contiguous(), dL[3], sc.shs, sc.sh_degree,
                                                 sc.campos, o[6], o[0], o[7], o[8], o[3], sem, False, _band=(k, 0),
                                                 _height=H, _width=W, _stage=1)
        return o, rec
    for _ in range(3):
        o, rec = step()
    torch.cuda.synchronize()
    lib.grpg_profile_begin()
    for _ in range(steps):
        step()
    buf = C.create_string_buffer(8192)
    lib.grpg_profile_end(buf, 8192)
    kern = {}
    for line in buf.value.decode().strip().splitlines():
        n, c, ms = line.split(":")
        kern[n] = round(float(ms) / steps, 4)
    print(json.dumps({"stride": k, "env": env, "num_rendered": int(o[0]), "kernels_ms": kern,
                      "checksum": [float(o[1].double().sum()), float(rec.double().abs().sum())]}))
