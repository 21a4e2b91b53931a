"""GPU parity of the sky cube-map kernels (grpg_sky_forward / backward / compose_rgb8 through
gaussianrpg_b200.sky_cubemap) with the numpy oracle, and of the fused epilogue with the separate kernels."""
import numpy as np
import pytest
import torch

import sky_cases
from gaussianrpg_b200 import image_utils, sky_cubemap
from oracle import sky_oracle as so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(sky_cases.cases().keys()))
def test_sky_forward_backward_vs_oracle(name, cuda_device):
    c = sky_cases.cases()[name]
    x = sky_cases.inputs(c)
    H, W = c["H"], c["W"]
    dev = cuda_device
    cube = x["cube"].to(dev).requires_grad_(True)
    acc = None if x["acc"] is None else x["acc"].to(dev)
    jit = None if x["jitter"] is None else x["jitter"].to(dev)
    sky = sky_cubemap.sky_color(cube, H, W, x["K"].to(dev), x["R"].to(dev), x["T"].to(dev), acc=acc,
                                white_background=x["white"], jitter=jit)
    (sky * x["dL"].to(dev)).sum().backward()
    args = (H, W, x["K"].numpy(), x["R"].numpy(), x["T"].numpy())
    kw = dict(acc=sky_cases.np_(x["acc"]), fill=1.0 if x["white"] else 0.0, jitter=sky_cases.np_(x["jitter"]))
    want, col, mask = so.sky_forward(sky_cases.np_(x["cube"]), *args, **kw)
    got = sky.detach().cpu().numpy()
    # float32 ray directions against float64: a tap weight moves by ~1e-6 * res; pixels whose footprint straddles a
    # texel boundary may pick the neighbouring quad, which changes nothing (the lookup is continuous)
    tol = 2e-5 * c["res"]
    assert np.abs(got - want).max() <= tol, np.abs(got - want).max()
    assert got.min() >= 0.0 and got.max() <= 1.0
    gwant = so.sky_backward(sky_cases.np_(x["cube"]), *args, sky_cases.np_(x["dL"]), **kw)
    ggot = cube.grad.cpu().numpy()
    # pixels whose colour sits within `tol` of the clamp bounds may pass or block their gradient differently
    near = ((np.abs(col) < tol) | (np.abs(col - 1) < tol)).any(-1) & mask
    slack = np.abs(sky_cases.np_(x["dL"])).max() * near.sum() + 1e-4 * np.abs(gwant).max()
    assert np.abs(ggot - gwant).max() <= 1e-4 * np.abs(gwant).max() + tol * np.abs(sky_cases.np_(x["dL"])).max() * 8 + slack


def test_fused_sky_epilogue_equals_the_separate_kernels(cuda_device):
    c = sky_cases.cases()["street_like"]
    x = sky_cases.inputs(c)
    H, W, dev = c["H"], c["W"], cuda_device
    g = torch.Generator().manual_seed(9)
    rgb = (torch.rand(3, H, W, generator=g) * 1.2 - 0.1).to(dev)
    acc = x["acc"].to(dev)
    cube = x["cube"].to(dev)
    K, R = x["K"].to(dev), x["R"].to(dev)
    sky = sky_cubemap.sky_color(cube, H, W, K, R, acc=acc)
    want8, wantf = image_utils.compose_rgb8(rgb, acc, sky, return_float=True)
    got8, gotf = sky_cubemap.sky_compose_rgb8(cube, H, W, K, R, rgb, acc, return_float=True)
    assert torch.equal(got8, want8) and torch.equal(gotf, wantf)
    # and the reference's expressions in stock torch ops (street_gaussian_renderer.py:336-346, simulator.py:313-314)
    ref = torch.clamp(rgb + sky * (1 - acc), 0.0, 1.0)
    ref8 = (ref.cpu().numpy().transpose(1, 2, 0) * 255).astype(np.uint8)
    assert np.array_equal(got8.cpu().numpy(), ref8)
    host = torch.empty(H, W, 3, dtype=torch.uint8).pin_memory()
    sky_cubemap.sky_compose_rgb8(cube, H, W, K, R, rgb, acc, out=host)
    torch.cuda.synchronize()
    assert np.array_equal(host.numpy(), ref8)


def test_sky_bench_size_and_errors(cuda_device):
    dev = cuda_device
    H, W, res = 1280, 1920, 1024
    g = torch.Generator().manual_seed(2)
    cube = torch.rand(6, res, res, 3, generator=g).to(dev).requires_grad_(True)
    K, R, T = (t.to(dev) for t in sky_cases.camera(H, W, 0.1, -0.02, 2083.09))
    acc = torch.rand(1, H, W, generator=g).to(dev)
    sky = sky_cubemap.sky_color(cube, H, W, K, R, T, acc=acc)
    sky.sum().backward()
    mask = (1 - acc[0]) > 1e-3
    assert bool((sky[:, ~mask] == 0).all()) and float(sky.max()) <= 1.0
    # every masked pixel spreads exactly one unit of weight per channel (cube values lie inside (0, 1): no clamping)
    assert abs(float(cube.grad.sum()) - 3.0 * int(mask.sum())) <= 1e-3 * 3.0 * int(mask.sum())
    with pytest.raises(RuntimeError):
        sky_cubemap.sky_color(torch.rand(6, 4, 4, 3), 8, 8, K.cpu(), R.cpu())            # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        sky_cubemap.sky_color(torch.rand(6, 4, 5, 3, device=dev), 8, 8, K, R)            # not square
