"""The loss oracle (oracle/loss_oracle.py) against golden vectors produced by the reference's own
lib/utils/loss_utils.py (tests/golden/make_loss_golden.py).  float64 oracle vs float32 reference: values to
2e-6; gradients to 1e-3 of the tensor's max (the bar of the rasterizer's gradients; the reference's own float32
E[x^2]-mu^2 cancellation reaches 2e-4 on piecewise-constant images) plus a floor of 1e-3 / #elements for
gradients that are analytically zero (identical images: the reference returns 5e-10 of rounding noise)."""
from pathlib import Path

import numpy as np
import pytest

import loss_cases
from oracle import loss_oracle

GOLDEN = Path(__file__).resolve().parent / "golden"


def grad_close(a, b, rtol=1e-3):
    return float(np.abs(a - b).max()) <= rtol * float(np.abs(b).max()) + 1e-3 / b.size


@pytest.mark.parametrize("name", list(loss_cases.cases().keys()))
def test_loss_oracle_vs_reference_golden(name):
    gold = np.load(GOLDEN / f"loss_{name}.npz")
    mask = gold["mask"] if "mask" in gold else None
    out = loss_oracle.l1_ssim(gold["img1"], gold["img2"], mask, lambda_dssim=loss_cases.LAMBDA_DSSIM)
    assert abs(out["ssim"] - float(gold["ssim"])) <= 2e-6
    if gold["img1"].ndim == 4:
        assert np.abs(out["ssim_per_image"] - gold["ssim_per_image"]).max() <= 2e-6
        return
    assert abs(out["l1"] - float(gold["l1"])) <= 2e-6
    assert abs(out["loss"] - float(gold["loss"])) <= 2e-6
    for k in ("grad_l1", "grad_ssim", "grad_loss"):
        assert grad_close(out[k], gold[k]), k


def test_window_matches_the_reference_float32_window():
    w = loss_oracle.window_1d()
    assert w.shape == (11,) and abs(w.sum() - 1.0) < 1e-6 and np.allclose(w, w[::-1])
    assert abs(w[5] - 0.26601174) < 1e-6  # gaussian(11, 1.5)[5] as float32
