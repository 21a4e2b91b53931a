"""GPU parity of the fused scene-graph compose (grpg_compose_forward/backward through gaussianrpg_b200.scene_compose)
with the goldens produced by the reference's own StreetGaussianModel getters + autograd, with the numpy oracle, and --
at the bench size -- with the reference formulation written in stock PyTorch ops on the same device."""
import numpy as np
import pytest
import torch

import compose_cases
from gaussianrpg_b200 import scene_compose
from oracle import compose_oracle

pytestmark = pytest.mark.gpu
PARAMS = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
OUT_TOL, GRAD_TOL = 2e-6, 2e-5  # relative to the tensor maximum (float32 rounding of exp / sigmoid / normalise)


def _sub(d, dev, grad=True):
    return scene_compose.SubModel(*(d[k].to(dev).clone().requires_grad_(grad) for k in PARAMS))


def _run(case, dev):
    bk = _sub(case["bkgd"], dev) if case["bkgd"] is not None else None
    actors = [_sub(a, dev) for a in case["actors"]]
    K = len(actors)
    rots = case["obj_rots"].to(dev).clone().requires_grad_(True) if K else None
    trans = case["obj_trans"].to(dev).clone().requires_grad_(True) if K else None
    idft = [scene_compose.idft_base(t, case["F"]) for t in case["times"]]
    out = scene_compose.compose_scene(bk, actors, rots, trans, idft, [f.to(dev) for f in case["flips"]])
    w = compose_cases.out_weights(case)
    loss = sum((getattr(out, k) * w[k].to(dev)).sum() for k in w)
    loss.backward()
    return out, bk, actors, rots, trans


def _close(a, b, tol, what):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size:
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
        assert err <= tol, (what, err)


@pytest.mark.parametrize("name", list(compose_cases.cases().keys()))
def test_compose_vs_reference_golden(name, cuda_device, golden_dir):
    case = compose_cases.cases()[name]
    gold = np.load(golden_dir / f"compose_{name}.npz")
    out, bk, actors, rots, trans = _run(case, cuda_device)
    for k in ("xyz", "rotation", "scaling", "opacity", "features"):
        _close(getattr(out, k), gold["out_" + k], OUT_TOL, k)
    subs = ([("bkgd", bk)] if bk is not None else []) + [(f"actor{k}", a) for k, a in enumerate(actors)]
    for tag, s in subs:
        for p in PARAMS:
            _close(getattr(s, p).grad, gold[f"grad_{tag}_{p}"], GRAD_TOL, f"{tag}.{p}")
    if actors:
        _close(rots.grad, gold["grad_obj_rots"], 1e-4, "obj_rots")  # float32 sums over the actor's Gaussians
        _close(trans.grad, gold["grad_obj_trans"], 1e-4, "obj_trans")


def test_eps_branch_matches_oracle(cuda_device):
    """A zero quaternion takes F.normalize's eps branch: output 0, gradient g / eps (oracle = reference formula)."""
    case = compose_cases.cases()["ragged"]
    out, bk, actors, _, _ = _run(case, cuda_device)
    assert float(out.rotation[2].abs().max()) == 0.0
    w = compose_cases.out_weights(case)["rotation"]
    assert torch.allclose(bk.rotation.grad[2].cpu(), w[2] / compose_oracle.EPS, rtol=1e-6)


def _torch_reference(bk, actors, rots, trans, idft, flips):
    """The reference's formulation (street_gaussian_model.py:295-384,438-453) with stock PyTorch ops."""
    import torch.nn.functional as F

    def qmul(a, b):
        aw, ax, ay, az = torch.unbind(a, -1)
        bw, bx, by, bz = torch.unbind(b, -1)
        return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                            aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)

    def q2m(r):
        q = r / torch.sqrt((r * r).sum(1))[:, None]
        R = torch.zeros((q.size(0), 3, 3), device=r.device)
        r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r_ * z); R[:, 0, 2] = 2 * (x * z + r_ * y)
        R[:, 1, 0] = 2 * (x * y + r_ * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r_ * x)
        R[:, 2, 0] = 2 * (x * z - r_ * y); R[:, 2, 1] = 2 * (y * z + r_ * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
        return R

    subs = ([bk] if bk is not None else []) + list(actors)
    scaling = torch.cat([torch.exp(s.scaling) for s in subs], 0)
    opacity = torch.cat([torch.sigmoid(s.opacity) for s in subs], 0)
    feats = [torch.cat((bk.features_dc, bk.features_rest), 1)] if bk is not None else []
    for a, b in zip(actors, idft):
        base = torch.tensor(b, device=a.xyz.device)
        feats.append(torch.cat([torch.sum(a.features_dc * base[..., None], dim=1, keepdim=True), a.features_rest], 1))
    xyzs = [bk.xyz] if bk is not None else []
    rot = [F.normalize(bk.rotation)] if bk is not None else []
    if actors:
        n = [a.xyz.shape[0] for a in actors]
        obj_rots = torch.cat([rots[k].expand(n[k], -1) for k in range(len(actors))], 0)
        obj_trans = torch.cat([trans[k].unsqueeze(0).expand(n[k], -1) for k in range(len(actors))], 0)
        flip = torch.cat([f.bool() for f in flips], 0)
        flipq = torch.tensor([[0.0, 0.0, 1.0, 0.0]], device=obj_rots.device)
        rl = torch.cat([F.normalize(a.rotation) for a in actors], 0).clone()
        rl[flip] = qmul(flipq, rl[flip])
        rot.append(F.normalize(qmul(obj_rots, rl)))
        xl = torch.cat([a.xyz for a in actors], 0).clone()
        xl[flip, scene_compose.FLIP_AXIS] *= -1
        xyzs.append(torch.einsum('bij, bj -> bi', q2m(obj_rots), xl) + obj_trans)
    return scene_compose.ComposedScene(torch.cat(xyzs, 0), torch.cat(rot, 0), scaling, opacity, torch.cat(feats, 0))


def test_bench_size_vs_torch_formulation(cuda_device):
    """1.84 M background + 8 x 20 k actors (BASELINE config #3's composition), outputs and gradients against the
    reference formulation in PyTorch on the same device."""
    dev = cuda_device
    case = compose_cases.make_case(11, 1_840_000, [20_000] * 8, M=4, F=5)
    bk = _sub(case["bkgd"], dev)
    actors = [_sub(a, dev) for a in case["actors"]]
    rots = case["obj_rots"].to(dev).requires_grad_(True)
    trans = case["obj_trans"].to(dev).requires_grad_(True)
    idft = [scene_compose.idft_base(t, case["F"]) for t in case["times"]]
    flips = [f.to(dev) for f in case["flips"]]
    w = {k: v.to(dev) for k, v in compose_cases.out_weights(case).items()}
    leaves = [t for s in [bk] + actors for t in s] + [rots, trans]

    def grads_of(out):
        loss = sum((getattr(out, k) * w[k]).sum() for k in w)
        gs = torch.autograd.grad(loss, leaves)
        return gs

    ours = scene_compose.compose_scene(bk, actors, rots, trans, idft, flips)
    want = _torch_reference(bk, actors, rots, trans, idft, flips)
    for k in ("xyz", "rotation", "scaling", "opacity", "features"):
        _close(getattr(ours, k), getattr(want, k).detach().cpu().numpy(), OUT_TOL, k)
    for i, (a, b) in enumerate(zip(grads_of(ours), grads_of(want))):
        tol = 2e-4 if i >= len(leaves) - 2 else GRAD_TOL  # pose: float32 sums over 20 k rows in both
        _close(a, b.detach().cpu().numpy(), tol, f"leaf {i}")


def test_argument_errors(cuda_device):
    case = compose_cases.cases()["no_flip"]
    bk = _sub(case["bkgd"], cuda_device, grad=False)
    actors = [_sub(a, cuda_device, grad=False) for a in case["actors"]]
    with pytest.raises(RuntimeError):
        scene_compose.compose_scene(None, [])
    with pytest.raises(RuntimeError):
        scene_compose.compose_scene(bk, actors)  # pose missing
    with pytest.raises(RuntimeError):
        scene_compose.compose_scene(_sub(case["bkgd"], torch.device("cpu"), grad=False))
    out = scene_compose.compose_scene(bk)  # background only, no pose needed
    assert out.xyz.shape == (100, 3) and torch.equal(out.xyz, bk.xyz)
