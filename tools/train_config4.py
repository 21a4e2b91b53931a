"""BASELINE config #4 at bench size: `--iters` training iterations (default 600) of the reference's loop on the
2 M-Gaussian street scene at 1920x1280 with densify / prune ON (every 100 iterations from `--densify-from`), both arms
(tools/train_harness.py), same iteration count, and the trajectory comparison: it/s over the timed window, Gaussian
counts after each densify, loss curves, PSNR(ours, reference) at checkpoints.  Prints one JSON line.

    python tools/train_config4.py [--iters 600] [--points 2000000] [--densify-from 100] [--time-from 100]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import train_harness as th  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=600)
ap.add_argument("--points", type=int, default=2_000_000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1280)
ap.add_argument("--cams", type=int, default=8)
ap.add_argument("--densify-from", type=int, default=100)
ap.add_argument("--time-from", type=int, default=100)
ap.add_argument("--skip-reference", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
import diff_gaussian_rasterization as ours_dgr  # noqa: E402
ref_dgr = None if args.skip_reference else th.load_reference_extension()
cams = th.Cameras(points=args.points, width=args.width, height=args.height, n_cams=args.cams, dev=dev)
th.make_ground_truth(cams, ours_dgr, dev)
densify_at = tuple(range(args.densify_from, args.iters + 1, th.CFG["densification_interval"]))
eval_at = tuple(sorted(set((1,) + densify_at + (args.iters,))))
out = {"config": {"workload": f"train loop, street scene {args.points} Gaussians as 1 background + 8 actors, "
                              f"{args.width}x{args.height}, {args.cams} cameras, L1 + 0.2 DSSIM, statistics, densify/prune every "
                              f"100 iterations from {args.densify_from}, Adam; thresholds of configs/example/waymo_train_002.yaml",
                  "iters": args.iters, "timed_window": [args.time_from, args.iters], "densify_at": list(densify_at)}}
runs = {}
for name, dgr in (("reference", ref_dgr), ("ours", ours_dgr)):
    if dgr is None:
        continue
    arm = th.Arm(name, dgr, cams, dev)
    p0 = arm.total()
    r = th.run_training(arm, args.iters, densify_at, eval_at=eval_at, time_from=args.time_from)
    runs[name] = r
    out[name] = {"ms_per_iter": round(r["ms_per_iter"], 3), "iters_per_s": round(1000.0 / r["ms_per_iter"], 2),
                 "P_start": p0, "P_end": arm.total(),
                 "P_after_densify": {str(it): sum(r["sizes"][it][0]) for it in densify_at},
                 "densify_events": {str(it): {k: sum(s[k] for s in r["sizes"][it][1]) for k in ("cloned", "split", "pruned")}
                                    for it in densify_at},
                 "loss_first": round(r["losses"][0], 6), "loss_last": round(r["losses"][-1], 6)}
    del arm
    torch.cuda.empty_cache()
if "reference" in runs:
    o, r = runs["ours"], runs["reference"]
    lo, lr = torch.tensor(o["losses"]), torch.tensor(r["losses"])
    out["speedup"] = round(out["reference"]["ms_per_iter"] / out["ours"]["ms_per_iter"], 2)
    out["loss_max_rel_diff"] = float(((lo - lr).abs() / lr.abs()).max())
    out["loss_max_rel_diff_before_first_densify"] = float(((lo - lr).abs() / lr.abs())[:max(1, args.densify_from - 1)].max())
    out["psnr_ours_vs_reference"] = {str(it): round(th.psnr(o["renders"][it], r["renders"][it]), 2) for it in eval_at}
    out["psnr_vs_ground_truth_end"] = {"ours": round(th.psnr(o["renders"][args.iters].clamp(0, 1), cams.gt[0]), 2),
                                       "reference": round(th.psnr(r["renders"][args.iters].clamp(0, 1), cams.gt[0]), 2)}
print(json.dumps(out))
