"""Drop-in for the image losses of the reference trainer (`lib/utils/loss_utils.py`), SURVEY 8(f) rank 2.

Same names and signatures -- `l1_loss(network_output, gt, mask=None)` (:21-37) and
`ssim(img1, img2, window_size=11, size_average=True, mask=None)` (:91-124) -- plus `l1_ssim_loss`, the combination
`train.py:116-118` builds from them, evaluated in ONE kernel launch (`grpg_l1_ssim`, include/grpg_loss.h) that
returns the value and keeps the gradient map for autograd.  The reference's version is 5 grouped 11x11
convolutions and ~20 elementwise kernels forward, and their transposes backward.

There is no CPU path: inputs must be CUDA tensors (as they are in train.py).  Gradients flow to the first image
only (the rendered one); the second is the ground truth.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(_lib.last_error())


def _planes(img: torch.Tensor):
    if img.ndim == 3:
        return 1, int(img.shape[0])
    if img.ndim == 4:
        return int(img.shape[0]), int(img.shape[1])
    raise RuntimeError("expected an image of shape (C, H, W) or (B, C, H, W)")


def _launch(img1, img2, mask, coef_l1: float, coef_ssim: float, want_grad: bool):
    """Returns (sums[planes,2] float64 on the device, grad or None)."""
    if not img1.is_cuda:
        raise RuntimeError("gaussianrpg_b200.loss_utils has no CPU path: inputs must be CUDA tensors")
    if img1.shape != img2.shape:
        raise RuntimeError(f"image shapes differ: {tuple(img1.shape)} vs {tuple(img2.shape)}")
    B, Cc = _planes(img1)
    H, W = int(img1.shape[-2]), int(img1.shape[-1])
    a = img1.detach().contiguous().float()
    b = img2.detach().to(img1.device).contiguous().float()
    m = None
    if mask is not None:
        if mask.numel() != B * H * W:
            raise RuntimeError(f"mask must have shape (1, H, W) per image, got {tuple(mask.shape)}")
        m = mask.to(device=img1.device).reshape(B, H, W).to(torch.uint8).contiguous()
    sums = torch.empty((B * Cc, 2), dtype=torch.float64, device=img1.device)
    grad = torch.empty_like(a) if want_grad else None
    args = _lib.L1SsimArgs()
    args.planes, args.planes_per_mask, args.height, args.width = B * Cc, Cc, H, W
    args.img1, args.img2 = a.data_ptr(), b.data_ptr()
    args.mask = m.data_ptr() if m is not None else None
    args.coef_l1, args.coef_ssim = float(coef_l1), float(coef_ssim)
    args.sums = sums.data_ptr()
    args.grad = grad.data_ptr() if grad is not None else None
    args.stream = _lib.current_stream_ptr(img1.device)
    with torch.cuda.device(img1.device):
        _check(_lib.load().grpg_l1_ssim(C.byref(args)))
    return sums, grad, m


class _L1Ssim(torch.autograd.Function):
    """value = w_l1 * L1 + w_ssim * SSIM (scalar); the kernel writes d(value)/d(img1) while computing it."""

    @staticmethod
    def forward(ctx, img1, img2, mask, w_l1: float, w_ssim: float):
        B, Cc = _planes(img1)
        H, W = int(img1.shape[-2]), int(img1.shape[-1])
        # mean denominators: loss_utils.py:35 (masked pixels x channels), :121 (all pixels)
        n_l1 = float(mask.sum().item()) * Cc if mask is not None else float(B * Cc * H * W)
        n_ssim = float(B * Cc * H * W)
        want_grad = bool(img1.requires_grad)
        sums, grad, _ = _launch(img1, img2, mask, w_l1 / max(n_l1, 1.0), w_ssim / n_ssim, want_grad)
        tot = sums.sum(0)
        l1 = tot[0] / n_l1 if n_l1 > 0 else torch.full((), float("nan"), dtype=torch.float64, device=img1.device)
        value = (w_l1 * l1 + w_ssim * tot[1] / n_ssim).to(img1.dtype)
        if want_grad:
            ctx.save_for_backward(grad)
        ctx.shape = img1.shape
        return value

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).reshape(ctx.shape), None, None, None, None


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """lib/utils/loss_utils.py:21-37 -- network_output, gt: (C, H, W); mask: (1, H, W) bool."""
    return _L1Ssim.apply(network_output, gt, mask, 1.0, 0.0)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True,
         mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """lib/utils/loss_utils.py:91-124.  `window_size` other than 11 is not implemented (the reference never passes
    one); `size_average=False` needs (B, C, H, W) input, as in the reference, and returns one value per image."""
    if window_size != 11:
        raise NotImplementedError("only the reference's window_size=11 is implemented")
    if size_average:
        return _L1Ssim.apply(img1, img2, mask, 0.0, 1.0)
    if img1.ndim != 4:
        raise IndexError("size_average=False needs (B, C, H, W) input (loss_utils.py:124)")
    if img1.requires_grad:
        raise NotImplementedError("size_average=False is evaluation-only here")
    B, Cc = _planes(img1)
    sums, _, _ = _launch(img1, img2, mask, 0.0, 0.0, False)
    per = sums[:, 1].reshape(B, Cc).sum(1) / float(Cc * img1.shape[-2] * img1.shape[-1])
    return per.to(img1.dtype)


def l1_ssim_loss(image: torch.Tensor, gt_image: torch.Tensor, lambda_dssim: float, lambda_l1: float = 1.0,
                 mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """train.py:116-118 in one launch:
    (1 - lambda_dssim) * lambda_l1 * l1_loss(image, gt, mask) + lambda_dssim * (1 - ssim(image, gt, mask=mask))."""
    return lambda_dssim + _L1Ssim.apply(image, gt_image, mask, (1.0 - lambda_dssim) * lambda_l1, -lambda_dssim)


def l2_loss(network_output, gt, mask=None):
    """lib/utils/loss_utils.py:39-55 (evaluation metric, not on the training path): plain tensor expression."""
    network_output, gt = network_output.permute(1, 2, 0), gt.permute(1, 2, 0)
    if mask is not None:
        network_output, gt = network_output[mask.squeeze(0)], gt[mask.squeeze(0)]
    return ((network_output - gt) ** 2).mean()


def psnr(img1, img2, mask=None):
    """lib/utils/loss_utils.py:61-78 (evaluation metric): plain tensor expression."""
    return 20 * torch.log10(1.0 / torch.sqrt(l2_loss(img1, img2, mask)))
