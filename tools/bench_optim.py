"""Measurement of the fused Adam step and densification statistics at BASELINE config #3's
composition -- 1.84 M background Gaussians + 8 actors x 20 k, M = 4, fourier_dim 5: ours (one launch each) vs the
reference's formulation (nine `torch.optim.Adam(..., eps=1e-15)` stepped in turn, gaussian_model.py:286-318,
street_gaussian_model.py:536-541; masked read-modify-write statistics, :555-578), CUDA events.  One JSON line;
roofline = algorithmic bytes (Adam: p, g, m, v read + p, m, v written = 28 B per element; statistics: radii + grad read
and, for visible rows, 4 floats read-modify-written) / kernel time vs the measured HBM peak."""
import ctypes as C
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import optim_cases
from gaussianrpg_b200 import optim, _lib
from test_optim_gpu import _training_setup, NAMES

dev = torch.device("cuda:0")
sizes = [1_840_000] + [20_000] * 8
g = torch.Generator().manual_seed(5)
ours, theirs, p_all = [], [], []
for n in sizes:
    case = dict(n=n, M=4, F=5 if n < 100_000 else 1, scale=3.0)
    a = {k: torch.nn.Parameter(torch.randn(*s, device=dev)) for k, s in optim_cases.SHAPES(n, 4, case["F"]).items()}
    b = {k: torch.nn.Parameter(v.detach().clone()) for k, v in a.items()}
    for d in (a, b):
        for p in d.values():
            p.grad = torch.randn_like(p) * 1e-3
    p_all.append(a)
    ours.append(_training_setup(a, case))
    theirs.append(torch.optim.Adam([dict(lr=grp["lr"], name=grp["name"], params=[b[k] for k in optim_cases.PARAMS if NAMES[k] == grp["name"]])
                                    for grp in ours[-1].param_groups], lr=0.0, eps=1e-15))
numel = sum(p.numel() for a in p_all for p in a.values())


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ref_step():
    for o in theirs:
        o.step()


res = {"adam_ours_ms": timeit(lambda: optim.fused_adam_step(ours)), "adam_ref_ms": timeit(ref_step)}

radii, grad, subs = optim_cases.stats_inputs(sizes, seed=3)
radii, grad = radii.to(dev), grad.to(dev)
stats = [optim.DensifyStats(*(s[k].to(dev) for k in ("max_radii2D", "xyz_gradient_accum", "denom"))) for s in subs]
want = [{k: s[k].to(dev).clone() for k in s} for s in subs]


def ref_stats():
    off, vis_all, rf = 0, radii > 0, radii.float()
    for w, n in zip(want, sizes):
        vis, gg = vis_all[off:off + n], grad[off:off + n]
        w["max_radii2D"][vis] = torch.max(w["max_radii2D"][vis], rf[off:off + n][vis])
        w["xyz_gradient_accum"][vis, 0:1] += torch.norm(gg[vis, :2], dim=-1, keepdim=True)
        w["xyz_gradient_accum"][vis, 1:2] += torch.norm(gg[vis, 2:], dim=-1, keepdim=True)
        w["denom"][vis] += 1
        off += n


res["stats_ours_ms"] = timeit(lambda: optim.update_densification_stats(stats, radii, grad))
res["stats_ref_ms"] = timeit(ref_stats)

lib = _lib.load()
lib.grpg_profile_begin()
for _ in range(20):
    optim.fused_adam_step(ours)
    optim.update_densification_stats(stats, radii, grad)
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
lib.grpg_profile_end(buf, len(buf))
prof = {}
for line in buf.value.decode().strip().splitlines():
    name, n, ms = line.rsplit(":", 2)
    prof[name] = float(ms) / int(n)
try:
    peak = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]); src = "MEASURED_PEAKS.json"
except Exception:
    peak, src = 6568.0, "fallback"
P = sum(sizes)
vis = int((radii > 0).sum())
adam_bytes = 28 * numel
stats_bytes = 16 * P + 32 * vis
out = {"op": "adam_step + densify_stats", "config": {"sub_models": len(sizes), "gaussians": P, "parameter_elements": numel, "visible": vis},
       **{k: round(v, 4) for k, v in res.items()},
       "speedup_adam": round(res["adam_ref_ms"] / res["adam_ours_ms"], 2),
       "speedup_stats": round(res["stats_ref_ms"] / res["stats_ours_ms"], 2),
       "kernels_ms": {k: round(v, 4) for k, v in prof.items()},
       "roofline": {"bound": "hbm", "peak": peak, "peak_source": src, "unit": "GB/s",
                    "adam": {"bytes": adam_bytes, "achieved": round(adam_bytes / prof["adam_step"] / 1e6, 1)},
                    "stats": {"bytes": stats_bytes, "achieved": round(stats_bytes / prof["densify_stats"] / 1e6, 1)}}}
for k in ("adam", "stats"):
    out["roofline"][k]["frac"] = round(out["roofline"][k]["achieved"] / peak, 3)
print(json.dumps(out))
