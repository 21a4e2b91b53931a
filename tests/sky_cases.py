"""Seeded inputs of the sky cube-map tests (shared by the CPU oracle tests and the GPU tests)."""
import math

import numpy as np
import torch


def camera(H, W, yaw=0.3, pitch=-0.1, f=None):
    """K, R (world-to-camera), T like the reference's cameras (fx = fy, principal point at the centre)."""
    f = f or 0.9 * W
    K = torch.tensor([[f, 0.0, W / 2], [0.0, f, H / 2], [0.0, 0.0, 1.0]])
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    Ry = torch.tensor([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    Rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, cp, -sp], [0.0, sp, cp]])
    return K, (Rx @ Ry).contiguous(), torch.tensor([0.3, -1.2, 2.0])


def cases():
    """name -> dict(H, W, res, yaw, pitch, f, acc?, white?, jitter?)"""
    return {
        "wide_lowres": dict(H=96, W=160, res=8, yaw=0.7, pitch=-0.6, f=40.0),      # very wide: sees 5 faces, corners, seams
        "street_like": dict(H=120, W=200, res=64, yaw=0.2, pitch=-0.05, f=170.0, acc=True),
        "up_white": dict(H=64, W=64, res=16, yaw=2.5, pitch=-1.3, f=30.0, acc=True, white=True),
        "jittered": dict(H=50, W=70, res=32, yaw=-1.1, pitch=0.4, f=60.0, jitter=True),
        "back": dict(H=33, W=47, res=5, yaw=3.1, pitch=0.9, f=20.0),               # ragged sizes, odd resolution
    }


def inputs(c, seed=0):
    g = torch.Generator().manual_seed(seed)
    H, W, res = c["H"], c["W"], c["res"]
    cube = torch.rand(6, res, res, 3, generator=g) * 1.4 - 0.2  # some texels outside [0, 1]: the clamp is exercised
    K, R, T = camera(H, W, c["yaw"], c["pitch"], c["f"])
    acc = None
    if c.get("acc"):
        acc = torch.rand(1, H, W, generator=g)
        acc[:, : H // 3] = 0.0          # open sky on top
        acc[:, -H // 4:] = 1.0          # fully covered at the bottom
    jitter = torch.rand(2, H, W, generator=g) if c.get("jitter") else None
    dL = torch.randn(3, H, W, generator=g)
    return dict(cube=cube, K=K, R=R, T=T, acc=acc, jitter=jitter, dL=dL, white=bool(c.get("white")))


def np_(t):
    return None if t is None else t.detach().cpu().numpy()
