"""Seeded inputs of the loss parity tests (shared by the golden generator, the CPU oracle tests and the GPU tests)."""
import torch

LAMBDA_DSSIM = 0.2  # configs/example/*.yaml, lib/config/config.py default


def _g(seed):
    return torch.Generator().manual_seed(seed)


def cases():
    out = {}
    g = _g(1)
    out["rand"] = (torch.rand(3, 40, 56, generator=g), torch.rand(3, 40, 56, generator=g), None)
    g = _g(2)
    a = torch.rand(3, 40, 56, generator=g)
    out["rand_mask"] = (a, torch.rand(3, 40, 56, generator=g), torch.rand(1, 40, 56, generator=g) > 0.3)
    g = _g(3)
    b = torch.rand(3, 37, 53, generator=g)
    out["ragged_close"] = ((b + 0.05 * torch.randn(3, 37, 53, generator=g)).clamp(0, 1), b, None)
    g = _g(4)
    out["tiny"] = (torch.rand(3, 7, 9, generator=g), torch.rand(3, 7, 9, generator=g), None)
    g = _g(5)
    a = torch.rand(3, 33, 47, generator=g)
    out["identical"] = (a, a.clone(), None)
    g = _g(6)
    out["gray"] = (torch.rand(1, 33, 20, generator=g), torch.rand(1, 33, 20, generator=g), torch.rand(1, 33, 20, generator=g) > 0.5)
    g = _g(7)  # piecewise-constant images: sigma ~ 0, the C1/C2 terms and cancellation in E[x^2]-mu^2 dominate
    blocks = torch.rand(3, 5, 6, generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)
    out["flat_blocks"] = (blocks, (blocks + 0.02 * (torch.rand(3, 40, 48, generator=g) - 0.5)).clamp(0, 1), None)
    g = _g(8)
    out["wide"] = (torch.rand(3, 21, 150, generator=g), torch.rand(3, 21, 150, generator=g), None)
    g = _g(9)
    out["batch"] = (torch.rand(2, 3, 24, 30, generator=g), torch.rand(2, 3, 24, 30, generator=g), None)
    return out
