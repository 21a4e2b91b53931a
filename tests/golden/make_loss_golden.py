"""Golden vectors for the L1 + SSIM loss, produced by IMPORTING the reference's own
`lib/utils/loss_utils.py` (pure PyTorch) on the CPU of the build container:

    python tests/golden/make_loss_golden.py          # needs /root/reference; writes tests/golden/loss_*.npz

The two imports of that module that need the rest of the product tree (`lib.config`, `lib.utils.img_utils`) are
stubbed; the functions under test (l1_loss :21-37, gaussian/create_window/ssim/_ssim :81-124) are untouched.
The committed .npz files are what the tests read -- /root/reference is not needed to run them.
"""
import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
import loss_cases  # noqa: E402


def load_reference():
    for name in ("lib.utils.img_utils", "lib.config"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["lib.utils.img_utils"].save_img_torch = lambda *a, **k: None
    sys.modules["lib.config"].cfg = None
    sys.path.insert(0, "/root/reference")
    return importlib.import_module("lib.utils.loss_utils")


def main():
    lu = load_reference()
    out_dir = ROOT / "tests" / "golden"
    for name, (a, b, mask) in loss_cases.cases().items():
        d = dict(img1=a.numpy(), img2=b.numpy())
        if mask is not None:
            d["mask"] = mask.numpy()
        if a.ndim == 3:
            for key, fn in (("l1", lambda x: lu.l1_loss(x, b, mask)),
                            ("ssim", lambda x: lu.ssim(x, b, mask=mask)),
                            ("loss", lambda x: (1.0 - loss_cases.LAMBDA_DSSIM) * lu.l1_loss(x, b, mask)
                             + loss_cases.LAMBDA_DSSIM * (1.0 - lu.ssim(x, b, mask=mask)))):  # train.py:118, lambda_l1 = 1
                x = a.clone().requires_grad_(True)
                v = fn(x)
                v.backward()
                d[key] = np.float64(v.item())
                d["grad_" + key] = x.grad.numpy()
        else:  # batched: the reference's size_average=False branch only works for 4-D input
            d["ssim"] = np.float64(lu.ssim(a, b).item())
            d["ssim_per_image"] = lu.ssim(a, b, size_average=False).numpy()
        np.savez_compressed(out_dir / f"loss_{name}.npz", **d)
        print(name, {k: (v if np.ndim(v) == 0 else v.shape) for k, v in d.items() if k not in ("img1", "img2", "mask")})


if __name__ == "__main__":
    main()
