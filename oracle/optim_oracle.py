"""CPU oracle (TEST INFRASTRUCTURE, never on the product path) for SURVEY 8f rank 3.

* `adam_step`: numpy float32 restatement of torch/optim/adam.py `_multi_tensor_adam` (non-capturable path, amsgrad off,
  weight_decay 0) -- the algorithm behind the reference's per-sub-model `torch.optim.Adam(l, lr=0.0, eps=1e-15)`
  (/root/reference/lib/models/gaussian_model.py:286-318).  torch (pinned here: 2.11.0) is the third-party dependency
  that holds the arithmetic; its published update is restated operation for operation:
      m.lerp_(g, 1-b1); v.mul_(b2); v.addcmul_(g, g, value=1-b2);
      denom = v.sqrt() / sqrt(1 - b2^t) + eps;  p.addcdiv_(m, denom, value=-lr / (1 - b1^t))
* `expon_lr`: get_expon_lr_func (lib/utils/general_utils.py:53-88), the xyz schedule of update_learning_rate (:320-325).
* `densify_stats`: StreetGaussianModel.set_max_radii2D + add_densification_stats (street_gaussian_model.py:555-578).
Pinned by tests/golden/optim_*.npz, produced by the reference's own GaussianModel.training_setup / update_optimizer and
StreetGaussianModel methods on the CPU (tests/golden/make_optim_golden.py).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """One update, float32 arrays in, new (p, m, v) out; `step` is the 1-based count AFTER the increment."""
    p, g, m, v = (np.asarray(x, dtype=F) for x in (p, g, m, v))
    m = m + F(1 - beta1) * (g - m)
    v = v * F(beta2)
    v = v + (F(1 - beta2) * g) * g
    step_size = F((lr / (1 - beta1 ** step)) * -1)
    bc2_sqrt = F((1 - beta2 ** step) ** 0.5)
    denom = np.sqrt(v) / bc2_sqrt + F(eps)
    p = p + step_size * (m / denom)
    return p, m, v


def expon_lr(step, lr_init, lr_final, lr_delay_mult=1.0, max_steps=1000000, lr_delay_steps=0):
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    delay = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1)) \
        if lr_delay_steps > 0 else 1.0
    t = np.clip(step / max_steps, 0, 1)
    return delay * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)


def densify_stats(radii, grad, subs):
    """subs: list of dict(max_radii2D [n], xyz_gradient_accum [n,2], denom [n,1]); returns updated copies."""
    out, off = [], 0
    radii = np.asarray(radii)
    grad = np.asarray(grad, dtype=F)
    for s in subs:
        n = s["denom"].shape[0]
        r, g = radii[off:off + n], grad[off:off + n]
        off += n
        vis = r > 0
        mr, acc, den = (np.array(s[k], dtype=F, copy=True) for k in ("max_radii2D", "xyz_gradient_accum", "denom"))
        mr[vis] = np.maximum(mr[vis], r[vis].astype(F))
        acc[vis, 0] += np.sqrt(g[vis, 0] * g[vis, 0] + g[vis, 1] * g[vis, 1])
        acc[vis, 1] += np.abs(g[vis, 2])
        den[vis] += F(1)
        out.append(dict(max_radii2D=mr, xyz_gradient_accum=acc, denom=den))
    return out
