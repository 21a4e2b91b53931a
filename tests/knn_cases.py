"""Seeded point clouds for the distCUDA2 tests (CPU oracle tests and GPU tests)."""
import numpy as np
import torch


def clouds(big: bool = False):
    g = torch.Generator().manual_seed(11)
    n = 200_000 if big else 6000
    out = {}
    out["uniform"] = torch.rand(n, 3, generator=g) * 20 - 10
    # street-like: dense ground plane near the camera, sparse far away, a few tight clusters (actors)
    ground = torch.stack([torch.rand(n // 2, generator=g) * 40 - 20, 1.8 + torch.randn(n // 2, generator=g) * 0.05,
                          torch.rand(n // 2, generator=g) ** 2 * 120], 1)
    cl = torch.randn(n // 2, 3, generator=g) * 0.3 + torch.randint(0, 8, (n // 2, 1), generator=g).float() * 7.0
    out["street_like"] = torch.cat([ground, cl], 0)
    # duplicates (zero distances), a plane (empty extent on one axis), a line
    d = torch.rand(n // 4, 3, generator=g)
    out["duplicates"] = torch.cat([d, d, d[: n // 8], torch.rand(n // 8, 3, generator=g)], 0)
    plane = torch.rand(n // 2, 3, generator=g)
    plane[:, 1] = 0.25
    out["plane"] = plane
    line = torch.zeros(max(n // 20, 50), 3)
    line[:, 0] = torch.rand(line.shape[0], generator=g) * 5
    out["line"] = line
    return out


def tiny():
    g = torch.Generator().manual_seed(12)
    return {f"P{p}": torch.rand(p, 3, generator=g) for p in (1, 2, 3, 4, 5, 129, 1025)}
