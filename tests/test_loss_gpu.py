"""GPU parity of the fused L1 + SSIM loss (gaussianrpg_b200/loss_utils.py -> grpg_l1_ssim) against

  * golden vectors produced by the reference's own lib/utils/loss_utils.py (tests/golden/loss_*.npz),
  * the float64 CPU oracle (oracle/loss_oracle.py),
  * at the training resolution (3 x 1280 x 1920) a plain PyTorch fp32 restatement of the reference formula on the
    same device, plus size-independent properties (ssim(x, x) = 1 with zero gradient, symmetry of the value).

Bars: loss values 1e-5 absolute (float32 on both sides); gradients 1e-3 of the tensor's max plus 1e-3/#elements
for analytically-zero gradients (see tests/test_loss_oracle_cpu.py).
"""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import loss_cases
from gaussianrpg_b200 import loss_utils
from oracle import loss_oracle

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
VAL_TOL = 1e-5


def grad_close(a, b, rtol=1e-3):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) <= rtol * float(np.abs(b).max()) + 1e-3 / b.size


def _run(fn, img1, *rest, **kw):
    x = img1.clone().requires_grad_(True)
    v = fn(x, *rest, **kw)
    v.backward()
    return float(v), x.grad.cpu().numpy()


@pytest.mark.parametrize("name", [n for n in loss_cases.cases() if n != "batch"])
def test_loss_vs_reference_golden_and_oracle(name, cuda_device):
    gold = np.load(GOLDEN / f"loss_{name}.npz")
    a = torch.from_numpy(gold["img1"]).to(cuda_device)
    b = torch.from_numpy(gold["img2"]).to(cuda_device)
    mask = torch.from_numpy(gold["mask"]).to(cuda_device) if "mask" in gold else None
    orc = loss_oracle.l1_ssim(gold["img1"], gold["img2"], gold["mask"] if "mask" in gold else None,
                              lambda_dssim=loss_cases.LAMBDA_DSSIM)
    for key, fn, kw in (("l1", loss_utils.l1_loss, dict(mask=mask)), ("ssim", loss_utils.ssim, dict(mask=mask)),
                        ("loss", loss_utils.l1_ssim_loss, dict(lambda_dssim=loss_cases.LAMBDA_DSSIM, mask=mask))):
        v, g = _run(fn, a, b, **kw)
        assert abs(v - float(gold[key])) <= VAL_TOL, (key, v, float(gold[key]))
        assert abs(v - orc[key]) <= VAL_TOL, (key, "oracle")
        assert grad_close(g, gold["grad_" + key]), key
        assert grad_close(g, orc["grad_" + key]), (key, "oracle")


def test_batched_input_and_per_image_values(cuda_device):
    gold = np.load(GOLDEN / "loss_batch.npz")
    a, b = torch.from_numpy(gold["img1"]).to(cuda_device), torch.from_numpy(gold["img2"]).to(cuda_device)
    assert abs(float(loss_utils.ssim(a, b)) - float(gold["ssim"])) <= VAL_TOL
    per = loss_utils.ssim(a, b, size_average=False).cpu().numpy()
    assert np.abs(per - gold["ssim_per_image"]).max() <= VAL_TOL
    with pytest.raises(IndexError):
        loss_utils.ssim(a[0], b[0], size_average=False)  # the reference fails the same way on 3-D input
    with pytest.raises(NotImplementedError):
        loss_utils.ssim(a, b, window_size=7)


def _torch_reference(img1, img2, lambda_dssim, mask=None):
    """The reference formula (loss_utils.py:21-37,81-124, train.py:118) with stock PyTorch ops, fp32."""
    w1 = torch.tensor(loss_oracle.window_1d(), dtype=torch.float32, device=img1.device)
    C = img1.shape[0]
    window = (w1[:, None] @ w1[None, :]).expand(C, 1, 11, 11).contiguous()
    if mask is not None:
        l1 = (img1.permute(1, 2, 0)[mask.squeeze(0)] - img2.permute(1, 2, 0)[mask.squeeze(0)]).abs().mean()
        x, y = torch.where(mask, img1, torch.zeros_like(img1)), torch.where(mask, img2, torch.zeros_like(img2))
    else:
        l1 = (img1 - img2).abs().mean()
        x, y = img1, img2
    conv = lambda t: F.conv2d(t, window, padding=5, groups=C)  # noqa: E731
    mu1, mu2 = conv(x), conv(y)
    s1, s2, s12 = conv(x * x) - mu1 * mu1, conv(y * y) - mu2 * mu2, conv(x * y) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim)


@pytest.mark.parametrize("masked", [False, True])
def test_training_resolution_against_torch_fp32(masked, cuda_device):
    g = torch.Generator().manual_seed(11)
    H, W = 1280, 1920
    gt = torch.rand(3, H, W, generator=g)
    gt = F.avg_pool2d(gt[None], 9, 1, 4)[0]  # smooth, image-like
    img = (gt + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    mask = (torch.rand(1, H, W, generator=g) > 0.1) if masked else None
    img, gt = img.to(cuda_device), gt.to(cuda_device)
    mask = mask.to(cuda_device) if masked else None
    v, gr = _run(loss_utils.l1_ssim_loss, img, gt, lambda_dssim=0.2, mask=mask)
    x = img.clone().requires_grad_(True)
    ref = _torch_reference(x, gt, 0.2, mask)
    ref.backward()
    assert abs(v - float(ref)) <= VAL_TOL
    assert grad_close(gr, x.grad.cpu().numpy())


def test_properties_at_training_resolution(cuda_device):
    g = torch.Generator().manual_seed(12)
    a = torch.rand(3, 1280, 1920, generator=g).to(cuda_device)
    b = torch.rand(3, 1280, 1920, generator=g).to(cuda_device)
    v, gr = _run(loss_utils.ssim, a, a.clone())
    assert abs(v - 1.0) <= 1e-6 and float(np.abs(gr).max()) <= 1e-3 / gr.size  # maximum of SSIM: zero gradient
    assert abs(float(loss_utils.ssim(a, b)) - float(loss_utils.ssim(b, a))) <= 1e-6  # the value is symmetric
    v1, g1 = _run(loss_utils.l1_loss, a, b)
    assert abs(v1 - float((a - b).abs().mean())) <= 1e-6
    assert np.array_equal(np.sign(g1), np.sign((a - b).cpu().numpy()))
    # ragged sizes: not multiples of the 32-pixel tile, smaller than the window
    for (h, w) in ((33, 65), (5, 3), (1, 1), (64, 31)):
        x, y = a[:, :h, :w].contiguous(), b[:, :h, :w].contiguous()
        o = loss_oracle.l1_ssim(x.cpu().numpy(), y.cpu().numpy(), None, lambda_dssim=0.2)
        v, gr = _run(loss_utils.l1_ssim_loss, x, y, lambda_dssim=0.2)
        assert abs(v - o["loss"]) <= VAL_TOL and grad_close(gr, o["grad_loss"]), (h, w)
