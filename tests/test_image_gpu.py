"""GPU parity of the fused image epilogue (grpg_compose_rgb8) with the reference's expressions
(oracle/image_oracle.py; the same lines run with torch on the CPU as a second opinion): bytes bit-exact."""
import numpy as np
import pytest
import torch

from gaussianrpg_b200 import image_utils
from oracle import image_oracle

pytestmark = pytest.mark.gpu


def _inputs(H, W, seed, spread=1.4):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(3, H, W, generator=g) * spread - 0.2  # some values outside [0, 1]
    acc = torch.rand(1, H, W, generator=g)
    sky = torch.rand(3, H, W, generator=g)
    return rgb, acc, sky


def _torch_cpu_reference(rgb, acc, sky):
    if sky is not None:
        rgb = rgb + sky * (1 - acc)                                             # street_gaussian_renderer.py:340
    rgb = torch.clamp(rgb, 0., 1.)                                              # :345
    return (rgb.detach().cpu().numpy().transpose(1, 2, 0) * 255).astype(np.uint8), rgb  # simulator.py:314


@pytest.mark.parametrize("H,W", [(1280, 1920), (64, 64), (37, 53), (1, 1), (3, 5), (16, 20)])
@pytest.mark.parametrize("with_sky", [False, True])
def test_bytes_bit_exact(H, W, with_sky, cuda_device):
    rgb, acc, sky = _inputs(H, W, 7 + H)
    want, wantf = image_oracle.compose_rgb8(rgb.numpy(), acc.numpy() if with_sky else None, sky.numpy() if with_sky else None)
    want_t, _ = _torch_cpu_reference(rgb, acc, sky if with_sky else None)
    assert np.array_equal(want, want_t)
    d = cuda_device
    out, outf = image_utils.compose_rgb8(rgb.to(d), acc.to(d) if with_sky else None, sky.to(d) if with_sky else None,
                                         return_float=True)
    assert np.array_equal(out.cpu().numpy(), want)
    assert np.array_equal(outf.cpu().numpy(), wantf)
    host = image_utils.to_host_rgb8(rgb.to(d), acc.to(d) if with_sky else None, sky.to(d) if with_sky else None)
    assert host.shape == (H, W, 3) and np.array_equal(host, want)  # device bytes + copy engine
    host = image_utils.to_host_rgb8(rgb.to(d), acc.to(d) if with_sky else None, sky.to(d) if with_sky else None, direct=True)
    assert np.array_equal(host, want)  # stored by the kernel into pinned memory


def test_exact_byte_boundaries(cuda_device):
    """Every float that lands exactly on or next to a k/255 boundary must truncate like numpy."""
    ks = torch.arange(0, 256, dtype=torch.float32) / 255.0
    vals = torch.cat([ks, torch.nextafter(ks, torch.tensor(2.0)), torch.nextafter(ks, torch.tensor(-1.0))])
    rgb = vals.repeat(3, 1).reshape(3, 1, -1).contiguous()
    want, _ = image_oracle.compose_rgb8(rgb.numpy())
    out, _ = image_utils.compose_rgb8(rgb.to(cuda_device))
    assert np.array_equal(out.cpu().numpy(), want)


def test_argument_errors(cuda_device):
    rgb = torch.rand(3, 8, 8, device=cuda_device)
    with pytest.raises(RuntimeError):
        image_utils.compose_rgb8(rgb.cpu())
    with pytest.raises(RuntimeError):
        image_utils.compose_rgb8(rgb, None, torch.rand(3, 8, 8, device=cuda_device))
    with pytest.raises(RuntimeError):
        image_utils.compose_rgb8(rgb, out=torch.empty(8, 8, 3, dtype=torch.uint8))  # pageable host memory
