"""Fused per-iteration parameter update: Adam for every sub-model in one launch, and the
densification statistics in one launch (include/grpg_optim.h).

Host-side mirror of the reference's optimiser plumbing:
* every `GaussianModel` owns `torch.optim.Adam(l, lr=0.0, eps=1e-15)` with seven named parameter groups
  (/root/reference/lib/models/gaussian_model.py:286-318) and `StreetGaussianModel.update_optimizer` steps them one
  after the other (street_gaussian_model.py:536-541).  `FusedAdam` IS a `torch.optim.Adam` (same constructor,
  `param_groups`, per-parameter `state` with `step` / `exp_avg` / `exp_avg_sq`), so the reference's learning-rate
  updates (:320-325), the optimiser-state surgery of densification (:394-470: `optimizer.state.get`, `del`,
  re-assignment) and `state_dict()` checkpoints work unchanged; only `step()` differs: one CUDA launch.
  `fused_adam_step([opt_bkgd, opt_obj_0, ...])` steps any number of them in ONE launch.
* `update_densification_stats` replaces `set_max_radii2D` + `add_densification_stats`
  (street_gaussian_model.py:555-578).
There is no CPU path.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import torch

from . import _lib


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"gaussianrpg_b200.optim has no CPU path: {what} must be a CUDA tensor")


def fused_adam_step(optimizers: Sequence[torch.optim.Adam]) -> int:
    """One launch for every parameter that has a gradient in every optimiser of the list (each a torch.optim.Adam
    or FusedAdam with amsgrad=False, weight_decay=0, maximize=False).  State layout and arithmetic follow
    torch/optim/adam.py (`_init_group`, `_multi_tensor_adam`, non-capturable path).  Returns the number of tensors
    updated."""
    rows = []
    dev = None
    for opt in optimizers:
        for group in opt.param_groups:
            if group.get("amsgrad") or group.get("weight_decay", 0) != 0 or group.get("maximize"):
                raise NotImplementedError("fused_adam_step: amsgrad / weight_decay / maximize are not used by the reference")
            beta1, beta2 = group["betas"]
            lr, eps = group["lr"], group["eps"]
            if isinstance(lr, torch.Tensor):
                lr = float(lr)
            for p in group["params"]:
                if p.grad is None:
                    continue  # torch skips parameters without a gradient: their moments do not decay
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                _require_cuda(p, "every parameter")
                state = opt.state[p]
                if len(state) == 0:  # adam.py `_init_group`
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                if p.numel() == 0:
                    continue
                step = float(state["step"])
                if not (p.is_contiguous() and p.dtype == torch.float32 and state["exp_avg"].is_contiguous()
                        and state["exp_avg_sq"].is_contiguous()):
                    raise RuntimeError("fused_adam_step: parameters and moments must be contiguous float32")
                g = p.grad if (p.grad.is_contiguous() and p.grad.dtype == torch.float32) else p.grad.float().contiguous()
                bias_correction1 = 1 - beta1 ** step
                bias_correction2 = 1 - beta2 ** step
                rows.append((p, g, state["exp_avg"], state["exp_avg_sq"], 1 - beta1, beta2, 1 - beta2,
                             (lr / bias_correction1) * -1, bias_correction2 ** 0.5, eps))
                dev = p.device if dev is None else dev
    if not rows:
        return 0
    lib = _lib.load()
    for begin in range(0, len(rows), 1024):
        part = rows[begin:begin + 1024]
        tab = (_lib.AdamTensor * len(part))()
        for t, (p, g, m, v, w1, b2, w2, ss, bc2, eps) in zip(tab, part):
            t.param, t.grad, t.exp_avg, t.exp_avg_sq, t.numel = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()
            t.one_minus_beta1, t.beta2, t.one_minus_beta2 = w1, b2, w2
            t.step_size, t.bias_correction2_sqrt, t.eps = ss, bc2, eps
        ws = torch.empty(int(lib.grpg_adam_workspace_bytes(len(part))), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            if lib.grpg_adam_step(tab, len(part), ws.data_ptr(), _lib.current_stream_ptr(dev)) != 0:
                raise RuntimeError(_lib.last_error())
    return len(rows)


class FusedAdam(torch.optim.Adam):
    """`torch.optim.Adam` whose `step()` is one CUDA launch over all parameter groups (see module docstring)."""

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        fused_adam_step([self])
        return loss


class DensifyStats(NamedTuple):
    """The three per-Gaussian statistics a reference `GaussianModel` keeps (gaussian_model.py:288-289,549)."""
    max_radii2D: torch.Tensor          # [n]
    xyz_gradient_accum: torch.Tensor   # [n,2]
    denom: torch.Tensor                # [n,1]


STATS_MAX_RADII, STATS_GRADIENTS = 1, 2  # GRPG_STATS_* (include/grpg_optim.h)


def update_densification_stats(stats: Sequence[DensifyStats], radii: Optional[torch.Tensor],
                               viewspace_point_tensor_grad: Optional[torch.Tensor],
                               visibility_filter: Optional[torch.Tensor] = None,
                               what: int = STATS_MAX_RADII | STATS_GRADIENTS) -> None:
    """`StreetGaussianModel.set_max_radii2D(radii, visibility_filter)` (what & STATS_MAX_RADII) and/or
    `add_densification_stats(viewspace_point_tensor, visibility_filter)` (what & STATS_GRADIENTS)
    (street_gaussian_model.py:555-578) for the sub-models of `graph_gaussian_range` in order, in place, in one launch.
    `radii` [P] (int32, the rasterizer's output), `viewspace_point_tensor_grad` [P,3]; P = total rows of the listed
    sub-models.  `visibility_filter` [P] bool: the mask the caller passes to the reference's methods; None means
    `radii > 0`, which is what the reference renderer builds (street_gaussian_renderer.py:268)."""
    ref_t = radii if radii is not None else visibility_filter
    if ref_t is None:
        raise RuntimeError("update_densification_stats: radii or visibility_filter is required")
    _require_cuda(ref_t, "radii / visibility_filter")
    lib = _lib.load()
    dev = ref_t.device
    P = int(ref_t.shape[0])
    if sum(int(s.denom.shape[0]) for s in stats) != P:
        raise RuntimeError("update_densification_stats: sub-model sizes must add up to len(radii); grad must be [P,3]")
    r = g = f = None
    if radii is not None:
        r = radii if (radii.dtype == torch.int32 and radii.is_contiguous()) else radii.to(torch.int32).contiguous()
    elif what & STATS_MAX_RADII:
        raise RuntimeError("update_densification_stats: radii is required for the max_radii2D update")
    if what & STATS_GRADIENTS:
        g = viewspace_point_tensor_grad
        if g is None or tuple(g.shape) != (P, 3):
            raise RuntimeError("update_densification_stats: sub-model sizes must add up to len(radii); grad must be [P,3]")
        g = g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous()
    if visibility_filter is not None:
        if tuple(visibility_filter.shape) != (P,) or visibility_filter.device != dev:
            raise RuntimeError("update_densification_stats: visibility_filter must be [P] on the radii's device")
        f = visibility_filter if visibility_filter.dtype in (torch.bool, torch.uint8) else visibility_filter != 0
        f = f.contiguous()
    tab = (_lib.StatsSubmodel * len(stats))()
    for t, s in zip(tab, stats):
        n = int(s.denom.shape[0])
        for x, shape in ((s.max_radii2D, (n,)), (s.xyz_gradient_accum, (n, 2)), (s.denom, (n, 1))):
            if tuple(x.shape) != shape or x.dtype != torch.float32 or not x.is_contiguous() or x.device != dev:
                raise RuntimeError("update_densification_stats: statistics must be contiguous float32 [n], [n,2], [n,1] on the radii's device")
        t.n = n
        if n:
            t.max_radii2D, t.xyz_gradient_accum, t.denom = s.max_radii2D.data_ptr(), s.xyz_gradient_accum.data_ptr(), s.denom.data_ptr()
    if P == 0:
        return
    ws = torch.empty(int(lib.grpg_stats_workspace_bytes(len(stats))), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if lib.grpg_densify_stats_ex(tab, len(stats), None if r is None else r.data_ptr(),
                                     None if g is None else g.data_ptr(), None if f is None else f.data_ptr(), int(what),
                                     ws.data_ptr(), _lib.current_stream_ptr(dev)) != 0:
            raise RuntimeError(_lib.last_error())
