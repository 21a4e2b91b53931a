"""A few forward+backward steps of the 2 M-Gaussian scene through the raw op (profiling target for ncu)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import cases
from gaussianrpg_b200 import synthetic, _C

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
dev = torch.device("cuda:0")
sc_cpu = synthetic.street_scene(P=P)
sc = sc_cpu.to(dev)
dL = [t.to(dev) for t in cases.loss_grads(sc_cpu)]
for _ in range(steps):
    fwd = cases.raw_forward(_C, sc)
    g = cases.raw_backward(_C, sc, fwd, dL)
torch.cuda.synchronize()
print("R", fwd[0])
