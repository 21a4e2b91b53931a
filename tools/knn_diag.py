"""Diagnostic: ours vs the brute-force oracle vs the reference simple-knn extension on a mid-sized cloud."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import importlib.util
import numpy as np, torch
import knn_cases
from oracle import oracle
spec = importlib.util.spec_from_file_location("build_ref", ROOT / "oracle" / "build_ref.py")
build_ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(build_ref)
from gaussianrpg_b200.simple_knn import distCUDA2
ref = build_ref.load_knn() if build_ref.knn_available() else None
dev = torch.device("cuda:0")
for name, pts in knn_cases.clouds(big=True).items():
    pts = pts[:30000]
    p = pts.to(dev)
    ours = distCUDA2(p).cpu().numpy()
    want = oracle.knn_mean_dist2(pts.numpy())
    line = f"{name}: ours vs oracle mismatches {(ours.view(np.uint32) != want.view(np.uint32)).sum()} of {len(want)}"
    if ref is not None:
        theirs = ref.distCUDA2(p).cpu().numpy()
        bad = theirs.view(np.uint32) != want.view(np.uint32)
        line += f"; reference vs oracle mismatches {bad.sum()}"
        if bad.any():
            i = np.nonzero(bad)[0][:5]
            line += f" e.g. idx {i.tolist()} ref {theirs[i].tolist()} oracle {want[i].tolist()} rel {(np.abs(theirs[i]-want[i])/want[i]).tolist()}"
        both = ours.view(np.uint32) != theirs.view(np.uint32)
        line += f"; ours vs reference mismatches {both.sum()}, max ulp {np.abs(ours.view(np.int32).astype(np.int64) - theirs.view(np.int32).astype(np.int64)).max()}"
    print(line, flush=True)
