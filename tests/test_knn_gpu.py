"""GPU parity of distCUDA2 (grpg_knn_mean_dist2 through gaussianrpg_b200.simple_knn and the drop-in package
`simple_knn._C`) with the CPU oracle and, when oracle/_ref holds it, with the UNMODIFIED reference extension: the
result is an exact 3-nearest-neighbour statistic with a fixed rounding, so the comparison is bit for bit."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import knn_cases
from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _ref():
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_ref", ROOT / "oracle" / "build_ref.py")
    build_ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build_ref)
    return build_ref.load_knn() if build_ref.knn_available() else None


def test_distcuda2_vs_oracle_bit_exact(cuda_device):
    from simple_knn._C import distCUDA2  # the import the reference uses (lib/models/gaussian_model.py:5)
    for name, pts in {**knn_cases.clouds(), **knn_cases.tiny()}.items():
        got = distCUDA2(pts.to(cuda_device)).cpu().numpy()
        want = oracle.knn_mean_dist2(pts.numpy())
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
    assert distCUDA2(torch.zeros(0, 3, device=cuda_device)).shape == (0,)
    with pytest.raises(RuntimeError):
        distCUDA2(torch.rand(5, 3))  # CPU tensor: no CPU path


def test_distcuda2_vs_reference_extension(cuda_device):
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref/simple_knn (reference extension) not built")
    from gaussianrpg_b200.simple_knn import distCUDA2
    for name, pts in {**knn_cases.clouds(big=True), **{k: v for k, v in knn_cases.tiny().items() if v.shape[0] >= 4}}.items():
        p = pts.to(cuda_device)
        got, want = distCUDA2(p), ref.distCUDA2(p)
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), name
    # what the reference does with it: the initial scales of create_from_pcd (gaussian_model.py:63-64)
    p = knn_cases.clouds(big=True)["street_like"].to(cuda_device)
    ours = torch.log(torch.sqrt(torch.clamp_min(distCUDA2(p), 0.0000001)))[..., None].repeat(1, 3)
    theirs = torch.log(torch.sqrt(torch.clamp_min(ref.distCUDA2(p), 0.0000001)))[..., None].repeat(1, 3)
    assert torch.equal(ours, theirs)


def test_distcuda2_bench_size_timing(cuda_device):
    """2 M points (the bench scene's cloud): runs, finite, and the time next to the reference extension's is printed."""
    from gaussianrpg_b200 import synthetic
    from gaussianrpg_b200.simple_knn import distCUDA2
    p = synthetic.street_scene().means3D.to(cuda_device)
    out = distCUDA2(p)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all()) and float(out.min()) >= 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); distCUDA2(p); e1.record(); torch.cuda.synchronize()
    msg = f"distCUDA2 2 M points: {e0.elapsed_time(e1):.2f} ms"
    ref = _ref()
    if ref is not None:
        want = ref.distCUDA2(p)
        e0.record(); ref.distCUDA2(p); e1.record(); torch.cuda.synchronize()
        msg += f" (reference extension: {e0.elapsed_time(e1):.2f} ms)"
        assert torch.equal(out.view(torch.int32), want.view(torch.int32))
    print(msg)
    out_dir = ROOT / "gpurun_out"
    if out_dir.is_dir():
        (out_dir / "knn_timing.txt").write_text(msg + "\n")
