"""Seeded inputs for the scene-graph compose op: sub-model parameters exactly as the reference
stores them (lib/models/gaussian_model.py:207-251 -- raw xyz, log-scales, raw quaternions, opacity logits, SH dc /
rest; actors carry `fourier_dim` dc rows, gaussian_model_actor.py:73-82) plus the per-actor pose that
`StreetGaussianModel.parse_camera` derives (street_gaussian_model.py:265-293)."""
from __future__ import annotations

import numpy as np
import torch

FLIP_AXIS = 1  # street_gaussian_model.py:58


def _sub(g, n, M, F):
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(xyz=r(n, 3) * 3.0, scaling=r(n, 3) * 0.6 - 2.0, rotation=r(n, 4), opacity=r(n, 1) * 1.5,
                features_dc=r(n, F, 3) * 0.5, features_rest=r(n, M - 1, 3) * 0.1)


def make_case(seed, n_bkgd, n_actors, M, F, flip_prob=0.5, unit_pose=False):
    g = torch.Generator().manual_seed(seed)
    case = dict(M=M, F=F, bkgd=_sub(g, n_bkgd, M, 1) if n_bkgd is not None else None, actors=[], times=[], flips=[])
    K = len(n_actors)
    for n in n_actors:
        case["actors"].append(_sub(g, n, M, F))
        case["times"].append(float(torch.rand((), generator=g)))  # fourier_scale * normalised frame
        case["flips"].append((torch.rand(n, generator=g) < flip_prob) if flip_prob > 0 else torch.zeros(n, dtype=torch.bool))
    q = torch.randn(K, 4, generator=g)
    if unit_pose:
        q = q / q.norm(dim=1, keepdim=True)
    else:
        q = q * (0.5 + torch.rand(K, 1, generator=g))  # quaternion_to_matrix normalises, the rotation product does not
    case["obj_rots"] = q
    case["obj_trans"] = torch.randn(K, 3, generator=g) * 10.0
    return case


def cases():
    c = {
        "street_small": make_case(1, 3000, [500, 777, 1], M=4, F=5),
        "actors_only": make_case(2, None, [333, 64], M=4, F=1),
        "bkgd_only": make_case(3, 1025, [], M=16, F=1),
        "ragged": make_case(4, 5, [1, 2, 1030], M=9, F=3, unit_pose=True),
        "no_flip": make_case(5, 100, [257, 255], M=4, F=5, flip_prob=0.0),
    }
    # a zero quaternion (F.normalize's eps branch) and a huge logit / log-scale
    c["ragged"]["bkgd"]["rotation"][2] = 0.0
    c["ragged"]["bkgd"]["opacity"][3] = 40.0
    c["ragged"]["actors"][2]["rotation"][7] = 0.0
    return c


def total(case):
    n = 0 if case["bkgd"] is None else case["bkgd"]["xyz"].shape[0]
    return n + sum(a["xyz"].shape[0] for a in case["actors"])


def out_weights(case, seed=99):
    """Seeded cotangents for (xyz, rotation, scaling, opacity, features): the loss is sum(w * out)."""
    g = torch.Generator().manual_seed(seed)
    P, M = total(case), case["M"]
    return dict(xyz=torch.randn(P, 3, generator=g), rotation=torch.randn(P, 4, generator=g),
                scaling=torch.randn(P, 3, generator=g), opacity=torch.randn(P, 1, generator=g),
                features=torch.randn(P, M, 3, generator=g))


def idft_base(time: float, dim: int) -> np.ndarray:
    """IDFT(time, dim)[0] of lib/utils/sh_utils.py:120-130 in float32 (cos(pi t k) for even k, sin(pi t (k+1)) for odd k)."""
    t = torch.tensor(time).view(-1, 1).float()
    idft = torch.zeros(1, dim)
    idx = torch.arange(dim)
    idft[:, idx[::2]] = torch.cos(torch.pi * t * idx[::2])
    idft[:, idx[1::2]] = torch.sin(torch.pi * t * (idx[1::2] + 1))
    return idft[0].numpy()
