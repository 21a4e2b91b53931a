"""Measurement of the fused scene-graph compose at BASELINE config #3's composition -- 1.84 M
background Gaussians + 8 actors x 20 k, SH degree 1 (M = 4), fourier_dim 5, flip_prob 0.5: ours (one launch forward,
one backward) vs the reference's formulation with stock PyTorch CUDA ops + autograd
(tests/test_compose_gpu.py::_torch_reference = street_gaussian_model.py:295-384,438-453), CUDA events.
Prints one JSON line; roofline = algorithmic bytes (every parameter read once, every output written once; backward:
cotangents + parameters read, parameter gradients written) / kernel time vs the measured HBM peak."""
import ctypes as C
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import compose_cases
from gaussianrpg_b200 import scene_compose, _lib
from test_compose_gpu import _torch_reference, _sub

dev = torch.device("cuda:0")
NB, NA, K, M, F = 1_840_000, 20_000, 8, 4, 5
case = compose_cases.make_case(11, NB, [NA] * K, M=M, F=F)
bk = _sub(case["bkgd"], dev)
actors = [_sub(a, dev) for a in case["actors"]]
rots = case["obj_rots"].to(dev).requires_grad_(True)
trans = case["obj_trans"].to(dev).requires_grad_(True)
idft = [scene_compose.idft_base(t, F) for t in case["times"]]
flips = [f.to(dev) for f in case["flips"]]
w = {k: v.to(dev) for k, v in compose_cases.out_weights(case).items()}
leaves = [t for s in [bk] + actors for t in s] + [rots, trans]
ours = lambda: scene_compose.compose_scene(bk, actors, rots, trans, idft, flips)  # noqa: E731
ref = lambda: _torch_reference(bk, actors, rots, trans, idft, flips)  # noqa: E731


def fwd_bwd(fn):
    out = fn()
    torch.autograd.backward([out.xyz, out.rotation, out.scaling, out.opacity, out.features],
                            [w["xyz"], w["rotation"], w["scaling"], w["opacity"], w["features"]])
    for t in leaves:
        t.grad = None


def fwd_only(fn):
    with torch.no_grad():
        fn()


def timeit(step, fn, n=30, warm=5):
    for _ in range(warm):
        step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(fn)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = dict(ours_fwd_ms=timeit(fwd_only, ours), ref_fwd_ms=timeit(fwd_only, ref),
           ours_fwd_bwd_ms=timeit(fwd_bwd, ours), ref_fwd_bwd_ms=timeit(fwd_bwd, ref))
lib = _lib.load()
lib.grpg_profile_begin()
for _ in range(20):
    fwd_bwd(ours)
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
P = NB + K * NA
par = 4 * (3 + 3 + 4 + 1 + 3 * (M - 1))          # bytes of parameters per Gaussian without dc
fwd_bytes = NB * (par + 12) + K * NA * (par + 12 * F) + P * 4 * (3 + 4 + 3 + 1 + 3 * M)
# backward: cotangents (same size as the outputs) + scaling/opacity/rotation (+ actors' xyz) read, all grads written
bwd_bytes = P * 4 * (3 + 4 + 3 + 1 + 3 * M) + P * 4 * (3 + 1 + 4) + K * NA * 12 + NB * (par + 12) + K * NA * (par + 12 * F)
out = {"op": "scene_compose", "config": {"background": NB, "actors": K, "per_actor": NA, "M": M, "fourier_dim": F},
       **{k: round(v, 4) for k, v in res.items()},
       "speedup_fwd": round(res["ref_fwd_ms"] / res["ours_fwd_ms"], 2),
       "speedup_fwd_bwd": round(res["ref_fwd_bwd_ms"] / res["ours_fwd_bwd_ms"], 2),
       "kernels_ms": {k: round(v, 4) for k, v in prof.items()},
       "roofline": {"bound": "hbm", "peak": peak, "peak_source": src, "unit": "GB/s",
                    "fwd": {"bytes": fwd_bytes, "achieved": round(fwd_bytes / prof["compose_fwd"] / 1e6, 1) if "compose_fwd" in prof else None},
                    "bwd": {"bytes": bwd_bytes, "achieved": round(bwd_bytes / prof["compose_bwd"] / 1e6, 1) if "compose_bwd" in prof else None}}}
for k in ("fwd", "bwd"):
    a = out["roofline"][k]["achieved"]
    out["roofline"][k]["frac"] = round(a / peak, 3) if a else None
print(json.dumps(out))
