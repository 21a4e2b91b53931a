"""Boundary conformance of the PyTorch surface on a GPU."""
from __future__ import annotations

import os

import pytest
import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from gaussianrpg_b200 import synthetic

pytestmark = pytest.mark.gpu


def _scene(dev, **kw):
    return synthetic.plumbing_scene(**kw).to(dev)


def test_five_tuple_and_shapes(cuda_device):
    sc = _scene(cuda_device, P=256, W=70, H=50, S=0)
    out = GaussianRasterizer(sc.settings())(means2D=None, **sc.raster_kwargs())
    color, radii, depth, alpha, sem = out
    assert color.shape == (3, 50, 70) and depth.shape == (1, 50, 70) and alpha.shape == (1, 50, 70)
    assert sem.shape == (0, 50, 70) and radii.shape == (256,) and radii.dtype == torch.int32


def test_semantics_s15_forward_and_backward(cuda_device):
    sc = _scene(cuda_device, P=300, W=64, H=48, S=15)
    kw = sc.raster_kwargs()
    kw["semantics"] = kw["semantics"].clone().requires_grad_(True)
    means2D = torch.zeros(300, 3, device=cuda_device, requires_grad=True)
    color, radii, depth, alpha, sem = GaussianRasterizer(sc.settings())(means2D=means2D, **kw)
    assert sem.shape == (15, 48, 64)
    (sem.sum() + color.sum()).backward()
    assert kw["semantics"].grad.shape == (300, 15) and means2D.grad.shape == (300, 3)
    assert float(means2D.grad[:, 2].min()) >= 0.0  # |grad| channel


def test_too_many_semantic_channels_in_backward_raises(cuda_device):
    sc = _scene(cuda_device, P=64, W=32, H=32, S=40)
    kw = sc.raster_kwargs()
    kw["opacities"] = kw["opacities"].clone().requires_grad_(True)
    color, radii, depth, alpha, sem = GaussianRasterizer(sc.settings())(means2D=None, **kw)
    assert sem.shape == (40, 32, 32)  # forward supports any S
    with pytest.raises(RuntimeError, match="semantic"):
        sem.sum().backward()


def test_argument_validation_messages(cuda_device):
    sc = _scene(cuda_device, P=16, W=32, H=32)
    r = GaussianRasterizer(sc.settings())
    kw = sc.raster_kwargs()
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means2D=None, **{**kw, "colors_precomp": torch.rand(16, 3, device=cuda_device)})
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means2D=None, **{**kw, "shs": None})
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means2D=None, **{**kw, "cov3D_precomp": torch.rand(16, 6, device=cuda_device)})
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means2D=None, **{**kw, "scales": None})
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        r(means2D=None, **{**kw, "means3D": torch.rand(16, 4, device=cuda_device)})


def test_empty_input(cuda_device):
    sc = _scene(cuda_device, P=16, W=40, H=24)
    kw = {k: (v[:0] if isinstance(v, torch.Tensor) else v) for k, v in sc.raster_kwargs().items()}
    kw["means3D"] = kw["means3D"].clone().requires_grad_(True)
    color, radii, depth, alpha, sem = GaussianRasterizer(sc.settings())(means2D=None, **kw)
    assert radii.numel() == 0 and float(color.abs().max()) == 0.0 and float(alpha.abs().max()) == 0.0
    color.sum().backward()
    assert kw["means3D"].grad.shape == (0, 3)


def test_background_only_enters_colour(cuda_device):
    sc = _scene(cuda_device, P=200, W=64, H=64)
    black = GaussianRasterizer(sc.settings())(means2D=None, **sc.raster_kwargs())
    sc_w = synthetic.plumbing_scene(P=200, W=64, H=64, white_bg=True).to(cuda_device)
    white = GaussianRasterizer(sc_w.settings())(means2D=None, **sc_w.raster_kwargs())
    assert torch.equal(black[2], white[2]) and torch.equal(black[3], white[3]) and torch.equal(black[1], white[1])
    assert not torch.equal(black[0], white[0])


def test_mark_visible_and_filter(cuda_device):
    sc = _scene(cuda_device, P=500, W=96, H=64)
    r = GaussianRasterizer(sc.settings())
    vis = r.markVisible(sc.means3D)
    assert vis.dtype == torch.bool and vis.shape == (500,)
    assert torch.equal(vis, sc.means3D[:, 2] > 0.2)  # identity view: depth = z
    radii_f, xy = r.visible_filter(sc.means3D, sc.scales, sc.rotations)
    color, radii, *_ = r(means2D=None, **sc.raster_kwargs())
    assert torch.equal(radii_f, radii) and xy.shape == (500, 2)


def test_debug_snapshot_on_failure(cuda_device, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    sc = _scene(cuda_device, P=32, W=32, H=32)
    st = sc.settings(debug=True)._replace(sh_degree=3)  # 4 coefficients cannot serve degree 3 -> the call fails
    with pytest.raises(RuntimeError):
        GaussianRasterizer(st)(means2D=None, **sc.raster_kwargs())
    assert os.path.exists("snapshot_fw.dump")


def test_non_default_stream_and_noncontiguous_inputs(cuda_device):
    sc = _scene(cuda_device, P=400, W=80, H=48)
    base = GaussianRasterizer(sc.settings())(means2D=None, **sc.raster_kwargs())
    kw = sc.raster_kwargs()
    kw["means3D"] = torch.cat([kw["means3D"], kw["means3D"]], 1)[:, :3]  # non-contiguous view
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        other = GaussianRasterizer(sc.settings())(means2D=None, **kw)
    s.synchronize()
    assert torch.equal(base[0], other[0]) and torch.equal(base[1], other[1])


def test_misaligned_views_take_the_plain_load_path(cuda_device):
    """preprocess stages 16-byte-aligned attribute slices with TMA bulk copies and reads quaternions as float4; a
    contiguous view that starts 4 bytes into its storage (every per-Gaussian input, rotations included) must give the
    same images, gradients and visible_filter results through the fallback loads."""
    sc = _scene(cuda_device, P=1500, W=96, H=64)

    def shifted(t):
        buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        buf[1:] = t.detach().reshape(-1)
        v = buf[1:].view(t.shape)
        assert v.data_ptr() % 16 != 0 and v.is_contiguous()
        return v

    def run(shift):
        kw = {}
        for k, v in sc.raster_kwargs().items():
            if torch.is_tensor(v) and k in ("means3D", "scales", "shs", "rotations", "opacities"):
                v = (shifted(v) if shift else v.detach().clone()).requires_grad_(True)
            kw[k] = v
        out = GaussianRasterizer(sc.settings())(means2D=None, **kw)
        (out[0].sum() + 0.1 * out[2].sum() + out[3].sum()).backward()
        return out, {k: v.grad for k, v in kw.items() if torch.is_tensor(v) and v.requires_grad}

    base, gbase = run(False)
    other, gother = run(True)
    for a, b in zip(base, other):
        assert torch.equal(a, b)
    assert set(gbase) == set(gother) and len(gbase) == 5
    for k in gbase:  # same kernels, same inputs: only the atomic order of the blend backward differs
        assert float((gbase[k] - gother[k]).abs().max()) <= 1e-3 * float(gbase[k].abs().max()), k
    r0, m0 = GaussianRasterizer(sc.settings()).visible_filter(sc.means3D, sc.scales, sc.rotations)
    r1, m1 = GaussianRasterizer(sc.settings()).visible_filter(shifted(sc.means3D), shifted(sc.scales), shifted(sc.rotations))
    assert torch.equal(r0, r1) and torch.equal(m0, m1)


def test_binning_workspace_guess_too_small_and_too_large(cuda_device):
    """The binning workspace is sized from the previous call with the same shape (grpg_forward); a guess that is too
    small must fall back to the exact two-stage path, a guess that is far too large must change nothing."""
    from gaussianrpg_b200 import _C
    import cases
    small = synthetic.plumbing_scene(P=3000, W=320, H=192, seed=3)
    big = synthetic.plumbing_scene(P=3000, W=320, H=192, seed=4)
    small.scales = small.scales * 0.2
    big.scales = big.scales * 4.0  # same shape, several times the instances (the guess carries 25 % slack)
    outs = {}
    for name, order in (("big_after_small", (small, big)), ("big_alone", (big,)), ("small_after_big", (big, small)),
                        ("small_alone", (small,))):
        _C._last_binned.clear()
        for sc in order:
            fwd = cases.raw_forward(_C, sc.to(cuda_device))
        outs[name] = fwd
    assert outs["big_after_small"][0] > 2 * outs["small_alone"][0]
    for a, b in (("big_after_small", "big_alone"), ("small_after_big", "small_alone")):
        assert outs[a][0] == outs[b][0]
        for i in (1, 2, 3, 5):
            assert torch.equal(outs[a][i], outs[b][i]), (a, i)
