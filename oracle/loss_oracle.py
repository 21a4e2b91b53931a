"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float64) of the reference's image loss, SURVEY 8f rank 2:

    l1_loss   lib/utils/loss_utils.py:21-37     mean |x - y| over the masked pixels (all channels)
    ssim      lib/utils/loss_utils.py:81-124    11-tap Gaussian window (sigma 1.5, :81-89), zero padding 5 (:105),
                                                 C1 = 0.01^2, C2 = 0.03^2 (:116-117), mask zeroes BOTH images (:96-98)
    loss      train.py:116-118                  (1 - lambda) * lambda_l1 * L1 + lambda * (1 - SSIM)

plus the gradient with respect to the first image, derived by hand (the reference relies on autograd).
Pinned by tests/golden/loss_*.npz, which were produced by importing the reference module itself
(tests/golden/make_loss_golden.py).  Only tests/ and the measurement tool's CPU baseline may import this file;
the product (gaussianrpg_b200/loss_utils.py) never does.
"""
from __future__ import annotations

import numpy as np

WINDOW, SIGMA = 11, 1.5
C1, C2 = 0.01 ** 2, 0.03 ** 2


# gaussian(11, 1.5) as the reference evaluates it (loss_utils.py:81-83: a float32 tensor divided by its float32
# sum); the literals are torch's result, window_1d() checks them against the formula to one float32 ulp
_WINDOW_F32 = np.array([float.fromhex(h) for h in (
    "0x1.0d956cp-10", "0x1.f1fe02p-8", "0x1.26eb18p-5", "0x1.bff0fep-4", "0x1.b43c3ep-3", "0x1.106560p-2",
    "0x1.b43c3ep-3", "0x1.bff0fep-4", "0x1.26eb18p-5", "0x1.f1fe02p-8", "0x1.0d956cp-10")])


def window_1d() -> np.ndarray:
    g = np.array([np.exp(-(x - WINDOW // 2) ** 2 / float(2 * SIGMA ** 2)) for x in range(WINDOW)])
    assert np.all(np.abs(_WINDOW_F32 - g / g.sum()) <= 2.0 ** -23 * _WINDOW_F32)
    return _WINDOW_F32.copy()


def _conv(planes: np.ndarray) -> np.ndarray:
    """Zero-padded 11x11 window = outer product of the 1-D window (loss_utils.py:85-89,105), per plane [...,H,W]."""
    w = window_1d()
    r = WINDOW // 2
    H, W = planes.shape[-2:]
    pad = np.zeros(planes.shape[:-2] + (H + 2 * r, W + 2 * r))
    pad[..., r:r + H, r:r + W] = planes
    tmp = sum(w[k] * pad[..., :, k:k + W] for k in range(WINDOW))
    return sum(w[k] * tmp[..., k:k + H, :] for k in range(WINDOW))


def ssim_map_and_grad(x: np.ndarray, y: np.ndarray):
    """Returns (ssim_map, d(sum of ssim_map)/dx) for planes [...,H,W] (already masked)."""
    mu1, mu2 = _conv(x), _conv(y)
    e11, e22, e12 = _conv(x * x), _conv(y * y), _conv(x * y)
    s1, s2, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    A1, A2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    B1, B2 = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    S = A1 * A2 / (B1 * B2)
    # partials of S with respect to the three window statistics that depend on x
    d_mu1 = (2 * mu2 * (A2 - A1)) / (B1 * B2) - S * (2 * mu1 / B1 - 2 * mu1 / B2)
    d_e11 = -S / B2
    d_e12 = 2 * A1 / (B1 * B2)
    grad = _conv(d_mu1) + 2 * x * _conv(d_e11) + y * _conv(d_e12)  # the window is symmetric
    return S, grad


def l1_ssim(img1, img2, mask=None, lambda_dssim=0.2, lambda_l1=1.0):
    """img: [C,H,W] (or [B,C,H,W]); mask: bool [1,H,W] (or [B,1,H,W]) or None.
    Returns dict(l1, ssim, loss, grad_l1, grad_ssim, grad_loss, ssim_per_image)."""
    x = np.asarray(img1, dtype=np.float64)
    y = np.asarray(img2, dtype=np.float64)
    batched = x.ndim == 4
    if not batched:
        x, y = x[None], y[None]
        mask = None if mask is None else np.asarray(mask)[None]
    B, C, H, W = x.shape
    m = np.ones((B, 1, H, W), dtype=bool) if mask is None else np.asarray(mask, dtype=bool).reshape(B, 1, H, W)
    mb = np.broadcast_to(m, x.shape)
    n_l1 = float(mb.sum())
    d = x - y
    l1 = float(np.abs(d[mb]).sum() / n_l1) if n_l1 else float("nan")
    grad_l1 = np.where(mb, np.sign(d), 0.0) / max(n_l1, 1.0)
    xm, ym = np.where(mb, x, 0.0), np.where(mb, y, 0.0)
    S, gS = ssim_map_and_grad(xm, ym)
    ssim = float(S.mean())
    grad_ssim = np.where(mb, gS, 0.0) / S.size
    out = dict(l1=l1, ssim=ssim, loss=(1.0 - lambda_dssim) * lambda_l1 * l1 + lambda_dssim * (1.0 - ssim),
               grad_l1=grad_l1, grad_ssim=grad_ssim,
               grad_loss=(1.0 - lambda_dssim) * lambda_l1 * grad_l1 - lambda_dssim * grad_ssim,
               ssim_per_image=S.reshape(B, -1).mean(1))
    if not batched:
        for k in ("grad_l1", "grad_ssim", "grad_loss"):
            out[k] = out[k][0]
    return out
