"""Seeded parity cases shared by the golden generator, the CPU oracle tests and the GPU tests."""
from __future__ import annotations

import numpy as np
import torch

from gaussianrpg_b200 import synthetic


def golden_cases():
    """name -> Scene.  Small enough that the reference outputs fit in a few hundred KB each."""
    cases = {}
    cases["plumbing_s3"] = synthetic.plumbing_scene(P=128, W=64, H=64, S=3, seed=0)
    cases["plumbing_white_deg3"] = synthetic.plumbing_scene(P=1024, W=100, H=76, S=0, white_bg=True, sh_degree=3, seed=1)
    cases["plumbing_deg2_ragged"] = synthetic.plumbing_scene(P=777, W=67, H=45, S=0, sh_degree=2, seed=2)
    cases["testscript_s15"] = synthetic.test_script_scene(P=96, W=310, H=94, S=15, seed=3)
    cases["street_small"] = synthetic.street_scene(P=40000, W=320, H=208, n_actors=2, actor_points=2000, seed=4)
    # precomputed colour / covariance branch (forward.cu:205-213,241)
    sc = synthetic.plumbing_scene(P=300, W=64, H=48, S=0, seed=5)
    g = torch.Generator().manual_seed(55)
    sc.colors_precomp = torch.rand(300, 3, generator=g)
    sc.shs = None
    A = torch.randn(300, 3, 3, generator=g) * 0.15
    cov = A @ A.transpose(1, 2)
    sc.cov3D_precomp = torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], 1).contiguous()
    sc.scales = None
    sc.rotations = None
    sc.name = "precomp"
    cases["precomp"] = sc
    cases["adversarial"] = adversarial_scene()
    return cases


def adversarial_scene(P: int = 3000, W: int = 208, H: int = 160, seed: int = 6):
    """Inputs chosen to stress the parts of our design that are NOT in the reference: needle-like Gaussians
    (ill-conditioned conics: the conservative culling must still never drop a contribution), splats hugging the
    near plane (radii of thousands of pixels), opacities straddling 1/255 and above 1, exact duplicates (depth
    ties -> stable order), and far-away points."""
    g = torch.Generator().manual_seed(seed)
    sc = synthetic.plumbing_scene(P=P, W=W, H=H, S=2, sh_degree=1, seed=seed)
    n = P // 6
    # needles: one axis 1000x longer than the others
    sc.scales[:n] = torch.exp(torch.rand(n, 3, generator=g) * 2 - 7)
    sc.scales[torch.arange(n), torch.randint(0, 3, (n,), generator=g)] *= 1000.0
    # near-plane splats
    sc.means3D[n:2 * n, 2] = 0.2 + torch.rand(n, generator=g) * 0.05
    sc.scales[n:2 * n] *= 0.2
    # opacity around the 1/255 threshold, exactly at it, and above one
    sc.opacities[2 * n:3 * n, 0] = (1.0 / 255.0) * (0.5 + torch.rand(n, generator=g) * 1.5)
    sc.opacities[3 * n:3 * n + 16, 0] = 1.0 / 255.0
    sc.opacities[3 * n + 16:3 * n + 64, 0] = 1.0 + torch.rand(48, generator=g)
    # duplicates (identical depth keys) and far points
    sc.means3D[4 * n:4 * n + 200] = sc.means3D[4 * n]
    sc.means3D[5 * n:5 * n + 100, 2] = 500.0 + torch.rand(100, generator=g) * 4000.0
    sc.name = "adversarial"
    return sc


def loss_grads(sc, seed=7):
    """Fixed pseudo-random dL/d(color, depth, alpha, semantic) so all three gradient inputs are exercised."""
    g = torch.Generator().manual_seed(seed)
    H, W = sc.height, sc.width
    S = 0 if sc.semantics is None else sc.semantics.shape[1]
    return (torch.randn(3, H, W, generator=g), torch.randn(1, H, W, generator=g) * 0.1,
            torch.randn(1, H, W, generator=g), torch.randn(S, H, W, generator=g))


def np_or_none(t):
    return None if t is None else t.detach().cpu().numpy()


def oracle_run(sc, with_backward=True):
    """Run the C oracle on a Scene; returns (pre, binned, img, grads|None)."""
    from oracle import oracle
    kw = dict(shs=np_or_none(sc.shs), sh_degree=sc.sh_degree, colors_precomp=np_or_none(sc.colors_precomp),
              scales=np_or_none(sc.scales), rotations=np_or_none(sc.rotations),
              cov3D_precomp=np_or_none(sc.cov3D_precomp), scale_modifier=sc.scale_modifier)
    pre, binned, img = oracle.forward(sc.means3D.numpy(), sc.opacities.numpy(), sc.viewmatrix.numpy(),
                                      sc.projmatrix.numpy(), sc.campos.numpy(), sc.width, sc.height, sc.tanfovx,
                                      sc.tanfovy, sc.bg.numpy(), semantics=np_or_none(sc.semantics), **kw)
    grads = None
    if with_backward:
        dc, dd, da, ds = [t.numpy() for t in loss_grads(sc)]
        grads = oracle.backward(sc.means3D.numpy(), sc.viewmatrix.numpy(), sc.projmatrix.numpy(), sc.campos.numpy(),
                                sc.width, sc.height, sc.tanfovx, sc.tanfovy, sc.bg.numpy(), pre, binned, img, dc, dd, da,
                                ds, semantics=np_or_none(sc.semantics), shs=np_or_none(sc.shs), sh_degree=sc.sh_degree,
                                scales=np_or_none(sc.scales), rotations=np_or_none(sc.rotations),
                                scale_modifier=sc.scale_modifier)
    return pre, binned, img, grads


def raw_forward(mod_C, sc, debug=False, **kw):
    """Call a `_C`-style module (ours or the reference's) with a Scene already on the GPU; `kw` = our
    keyword-only extras (e.g. `_reference_binning=True`)."""
    E = torch.Tensor([])
    P = sc.means3D.shape[0]
    sem = sc.semantics if sc.semantics is not None else torch.zeros(P, 0, device=sc.means3D.device)
    opt = lambda t: E if t is None else t  # noqa: E731
    return mod_C.rasterize_gaussians(sc.bg, sc.means3D, opt(sc.colors_precomp), sem, sc.opacities, opt(sc.scales),
                                     opt(sc.rotations), sc.scale_modifier, opt(sc.cov3D_precomp), sc.viewmatrix,
                                     sc.projmatrix, sc.tanfovx, sc.tanfovy, sc.height, sc.width, opt(sc.shs),
                                     sc.sh_degree, sc.campos, False, debug, **kw)


def raw_backward(mod_C, sc, fwd, dL, debug=False):
    E = torch.Tensor([])
    P = sc.means3D.shape[0]
    sem = sc.semantics if sc.semantics is not None else torch.zeros(P, 0, device=sc.means3D.device)
    opt = lambda t: E if t is None else t  # noqa: E731
    R, color, depth, alpha, semantic, radii, geom, binning, img = fwd
    dc, dd, da, ds = dL
    return mod_C.rasterize_gaussians_backward(sc.bg, sc.means3D, radii, opt(sc.colors_precomp), opt(sc.scales),
                                              opt(sc.rotations), sc.scale_modifier, opt(sc.cov3D_precomp),
                                              sc.viewmatrix, sc.projmatrix, sc.tanfovx, sc.tanfovy, dc, dd, da, ds,
                                              opt(sc.shs), sc.sh_degree, sc.campos, geom, R, binning, img, alpha, sem,
                                              debug)


GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations", "dL_dsemantics"]


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max |b| -- the per-tensor relative error of SURVEY 8(d)."""
    if a.size == 0:
        return 0.0
    den = float(np.abs(b).max())
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / (den + 1e-30)


def row_err(a: np.ndarray, b: np.ndarray, floor_frac: float = 1e-6) -> np.ndarray:
    """Per-Gaussian relative error (the per-element check with an absolute floor of SURVEY 8(d)): for every row i,
    max_j |a_ij - b_ij| / (max_j |b_ij| + floor) with floor = floor_frac * max |b|.  Unlike `rel_err` it does not let
    the few large-gradient Gaussians hide errors on the many small-gradient ones (far Gaussians, dL_drotations):
    every row whose gradient is more than a millionth of the tensor's largest is checked against its own scale."""
    if a.size == 0:
        return np.zeros(0)
    a2 = a.reshape(a.shape[0], -1).astype(np.float64)
    b2 = b.reshape(b.shape[0], -1).astype(np.float64)
    floor = floor_frac * float(np.abs(b2).max()) + 1e-30
    return np.abs(a2 - b2).max(axis=1) / (np.abs(b2).max(axis=1) + floor)


def row_violations(a: np.ndarray, b: np.ndarray, tol: float = 1e-3) -> float:
    """fraction of rows whose `row_err` exceeds tol"""
    e = row_err(a, b)
    return float((e > tol).mean()) if e.size else 0.0
