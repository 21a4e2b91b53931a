"""The image-epilogue oracle (numpy) against the reference's own expressions executed with torch on the CPU."""
import numpy as np
import pytest
import torch

from oracle import image_oracle


@pytest.mark.parametrize("with_sky", [False, True])
def test_numpy_restatement_equals_torch_cpu_expressions(with_sky):
    g = torch.Generator().manual_seed(3)
    rgb = torch.rand(3, 33, 47, generator=g) * 1.6 - 0.3
    acc, sky = torch.rand(1, 33, 47, generator=g), torch.rand(3, 33, 47, generator=g)
    r = rgb + sky * (1 - acc) if with_sky else rgb          # street_gaussian_renderer.py:340
    r = torch.clamp(r, 0., 1.)                               # :345
    want = (r.detach().cpu().numpy().transpose(1, 2, 0) * 255).astype(np.uint8)  # simulator.py:314
    got, gotf = image_oracle.compose_rgb8(rgb.numpy(), acc.numpy() if with_sky else None, sky.numpy() if with_sky else None)
    assert np.array_equal(got, want) and np.array_equal(gotf, r.numpy())
