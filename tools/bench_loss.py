"""Measurement of the fused L1 + SSIM loss at the training resolution, 3 x 1280 x 1920:
ours (one launch: value + gradient) vs the reference's algorithm run with stock PyTorch CUDA ops
(tests/test_loss_gpu.py::_torch_reference = lib/utils/loss_utils.py + autograd), forward + backward, CUDA events.
Prints one JSON line; roofline = algorithmic bytes (2 image reads + 1 gradient write) / kernel time vs HBM peak."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from gaussianrpg_b200 import loss_utils, _lib
from test_loss_gpu import _torch_reference

dev = torch.device("cuda:0")
H, W = 1280, 1920
g = torch.Generator().manual_seed(0)
gt = torch.rand(3, H, W, generator=g).to(dev)
img = (gt + 0.1 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1)


def step(fn):
    x = img.clone().requires_grad_(True)
    fn(x).backward()
    return x.grad


def timeit(fn, n=50, warm=10):
    for _ in range(warm):
        step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(fn)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ours = lambda x: loss_utils.l1_ssim_loss(x, gt, 0.2)  # noqa: E731
ref = lambda x: _torch_reference(x, gt, 0.2)  # noqa: E731
ms_ours, ms_ref = timeit(ours), timeit(ref)
# kernel-only time through the library's profiler
import ctypes as C
lib = _lib.load()
lib.grpg_profile_begin()
for _ in range(20):
    step(ours)
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
lib.grpg_profile_end(buf, len(buf))
prof = {}
for line in buf.value.decode().strip().splitlines():  # "name:launches:ms_total"
    name, n, ms = line.rsplit(":", 2)
    prof[name] = {"launches": int(n), "ms_total": float(ms)}
k_ms = prof["l1_ssim"]["ms_total"] / prof["l1_ssim"]["launches"] if "l1_ssim" in prof else None
peaks = {}
try:
    peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
except Exception:
    pass
peak = float(peaks.get("hbm_gbs", 6568.0)) if isinstance(peaks, dict) else 6568.0
alg = 3 * 3 * H * W * 4
out = {"workload": "L1 + 0.2 * DSSIM loss, value + gradient, 3x1280x1920 fp32", "ours_ms_fwd_bwd": ms_ours,
       "torch_reference_ms_fwd_bwd": ms_ref, "speedup": ms_ref / ms_ours, "kernel_ms": k_ms, "profile": prof,
       "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "peak": peak, "unit": "GB/s",
                    "achieved": (alg / (k_ms * 1e-3) / 1e9) if k_ms else None,
                    "frac": (alg / (k_ms * 1e-3) / 1e9 / peak) if k_ms else None}}
print(json.dumps(out))
