"""Determinism / validity stress of the binning stages on the 2 M scene."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import cases
from gaussianrpg_b200 import synthetic, _C, debug

dev = torch.device("cuda:0")
sc_cpu = synthetic.street_scene()
sc = sc_cpu.to(dev)
P, W, H = sc.means3D.shape[0], sc.width, sc.height
ref = None
bad = 0
Rs = {}
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    fwd = cases.raw_forward(_C, sc)
    R = fwd[0]
    Rs[R] = Rs.get(R, 0) + 1
    p = debug.parse_buffers(P, R, W, H, fwd[6], fwd[7], fwd[8])
    perm_ok = (lambda a_, b_: a_.numel() == b_.numel() and bool((a_ == b_).all()))(p["sorted_idx"].long().sort().values, (p["tiles_touched"] != 0).nonzero().flatten())
    tt = int(p["tiles_touched"].long().sum())
    keys = p["point_list_keys"]
    sorted_ok = bool((keys[1:] >= keys[:-1]).all())
    if ref is None:
        ref = (R, p["point_list"].clone(), fwd[1].clone())
    same = R == ref[0] and torch.equal(p["point_list"], ref[1]) and torch.equal(fwd[1], ref[2])
    if not (perm_ok and tt == R and sorted_ok and same):
        bad += 1
        print(f"iter {it}: R={R} sum_tiles={tt} perm_ok={perm_ok} sorted_ok={sorted_ok} same_as_first={same}", flush=True)
print("R histogram:", Rs, "bad iterations:", bad)
