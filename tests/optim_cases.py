"""Seeded inputs for the fused Adam step and the densification statistics."""
from __future__ import annotations

import types

import torch

# lib/config/config.py:48-57 (cfg.optim defaults); semantic_lr is read by training_setup (gaussian_model.py:299)
OPTIM_CFG = dict(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
                 position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
                 semantic_lr=0.01, percent_dense=0.01, percent_big_ws=0.1)
PARAMS = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation", "semantic")  # training_setup's group order
SHAPES = lambda n, M, F: dict(xyz=(n, 3), features_dc=(n, F, 3), features_rest=(n, M - 1, 3), opacity=(n, 1),  # noqa: E731
                              scaling=(n, 3), rotation=(n, 4), semantic=(n, 0))


class _Cfg(types.SimpleNamespace):
    """attribute access plus `.get(key, default)`, like the reference's yacs CfgNode"""

    def get(self, key, default=None):
        return getattr(self, key, default)


def optim_namespace():
    return _Cfg(**OPTIM_CFG)


def adam_cases():
    """name -> (n, M, F, spatial_lr_scale, first iteration, steps, set of (step, param) whose gradient is None)."""
    return {
        "bkgd": dict(n=1000, M=4, F=1, scale=7.3, it0=1, steps=6, missing={(2, "xyz"), (4, "opacity")}),
        "actor": dict(n=257, M=4, F=5, scale=1.9, it0=5000, steps=4, missing=set()),
        "tiny_deg3": dict(n=3, M=16, F=1, scale=0.5, it0=29990, steps=12, missing={(0, "rotation")}),
    }


def adam_params(case, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(*s, generator=g) for k, s in SHAPES(case["n"], case["M"], case["F"]).items()}


def adam_grads(case, step, seed=0):
    """Gradients of one step: a spread of magnitudes (1e-7 .. 1e-1) and exact zeros (invisible Gaussians)."""
    g = torch.Generator().manual_seed(1000 * (seed + 1) + step)
    out = {}
    for k, s in SHAPES(case["n"], case["M"], case["F"]).items():
        if (step, k) in case["missing"]:
            out[k] = None
            continue
        t = torch.randn(*s, generator=g) * torch.pow(10.0, torch.randint(-7, 0, s, generator=g).float())
        rows = torch.rand(s[0], generator=g) < 0.3
        t[rows] = 0.0
        out[k] = t
    return out


def stats_cases():
    """name -> list of sub-model sizes (background first)."""
    return {"street": [3000, 500, 777, 1], "one": [1025], "ragged": [5, 1, 2, 1030, 0, 7]}


def stats_inputs(sizes, seed=0):
    g = torch.Generator().manual_seed(seed)
    P = sum(sizes)
    radii = torch.randint(-1, 40, (P,), generator=g, dtype=torch.int32).clamp_min(0)  # ~5 % invisible
    radii[torch.rand(P, generator=g) < 0.3] = 0
    grad = torch.randn(P, 3, generator=g) * 1e-3
    subs = [dict(max_radii2D=torch.rand(n, generator=g) * 30, xyz_gradient_accum=torch.rand(n, 2, generator=g),
                 denom=torch.randint(0, 50, (n, 1), generator=g).float()) for n in sizes]
    return radii, grad, subs
