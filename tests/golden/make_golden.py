"""Generate golden vectors by RUNNING THE REFERENCE extension (oracle/_ref) on a GPU.

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/             # then commit them

Inputs are not stored: they are regenerated bit-identically from the seeds in tests/cases.py.
Stored per case: every observable output of the reference forward/backward plus the index
arrays parsed out of its opaque buffers (tests/refparse.py).
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))

import numpy as np
import torch

import build_ref
import cases
from refparse import parse_reference


def main():
    ref = build_ref.load()
    dev = torch.device("cuda:0")
    out_dir = ROOT / "gpurun_out" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    for name, sc_cpu in cases.golden_cases().items():
        sc = sc_cpu.to(dev)
        P, W, H = sc.means3D.shape[0], sc.width, sc.height
        fwd = cases.raw_forward(ref._C, sc)
        torch.cuda.synchronize()
        R = fwd[0]
        parsed = parse_reference(P, R, W, H, fwd[6], fwd[7], fwd[8])
        dL = [t.to(dev) for t in cases.loss_grads(sc_cpu)]
        grads = cases.raw_backward(ref._C, sc, fwd, dL)
        torch.cuda.synchronize()
        vis = (fwd[5] > 0)
        d = dict(R=np.int64(R), color=fwd[1].cpu().numpy(), depth=fwd[2].cpu().numpy(), alpha=fwd[3].cpu().numpy(),
                 semantic=fwd[4].cpu().numpy(), radii=fwd[5].cpu().numpy())
        # per-Gaussian intermediates are only defined for visible Gaussians in the reference (uninitialised otherwise)
        for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped"):
            v = parsed[k].clone()
            v[~vis] = 0
            d[k] = v.cpu().numpy()
        d["tiles_touched"] = parsed["tiles_touched"].cpu().numpy()
        d["n_contrib"] = parsed["n_contrib"].cpu().numpy()
        d["ranges"] = parsed["ranges"].cpu().numpy()
        if R > 0:
            d["point_list"] = parsed["point_list"].cpu().numpy()
            d["point_list_keys"] = parsed["point_list_keys"].cpu().numpy()
        for n, g in zip(cases.GRAD_NAMES, grads):
            d[n] = g.cpu().numpy()
        # The reference's backward sums with float atomics in scheduling order: record its own run-to-run spread
        # so that the tests can tell a wrong gradient from an ill-conditioned one (tests use max(1e-3, 3 x spread)).
        spread = {n: 0.0 for n in cases.GRAD_NAMES}
        for _ in range(4):
            again = cases.raw_backward(ref._C, sc, fwd, dL)
            torch.cuda.synchronize()
            for n, g in zip(cases.GRAD_NAMES, again):
                if g.numel():
                    spread[n] = max(spread[n], cases.rel_err(g.cpu().numpy(), d[n]))
        for n, v in spread.items():
            d["spread_" + n] = np.float64(v)
        np.savez_compressed(out_dir / f"{name}.npz", **d)
        print(name, "P", P, "R", R, "V", int(vis.sum()), "max grad spread %.2e" % max(spread.values()), flush=True)


if __name__ == "__main__":
    main()
