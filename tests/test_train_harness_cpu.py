"""Host logic of the config #4 harness (tools/train_harness.py) on the CPU: the lock-step runner and the forced
densify decisions, driven through a differentiable stand-in for the rasterizer (no CUDA).  Both arms use the
reference formulation (stock PyTorch compose / loss / Adam): the point is the bookkeeping, not the operators."""
import sys
from collections import namedtuple
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))


class _FakeDgr:
    GaussianRasterizationSettings = namedtuple(
        "GaussianRasterizationSettings", "image_height image_width tanfovx tanfovy bg scale_modifier viewmatrix "
        "projmatrix sh_degree campos prefiltered debug")

    class GaussianRasterizer:
        def __init__(self, raster_settings):
            self.s = raster_settings

        def __call__(self, means3D, means2D, opacities, shs, scales, rotations):
            P, H, W = means3D.shape[0], self.s.image_height, self.s.image_width
            col = (torch.sigmoid(opacities) * shs[:, 0, :]).sum(0) / P
            geo = (means3D.sum() + scales.exp().sum() + rotations.sum()) / P
            m2 = (means2D * means3D.detach().abs()).sum() / P  # all three columns: the third feeds the abs-gradient statistic
            img = (col + 0.01 * geo + m2)[:, None, None] * torch.ones(3, H, W)
            return img, torch.ones(P, dtype=torch.int32), img[:1], img[:1], None


def _arms(th, cams, cfg):
    return [th.Arm(n, _FakeDgr, cams, torch.device("cpu"), cfg=cfg, bkgd_extent=3.0, actor_extent=2.0)
            for n in ("reference", "reference-b")]


def test_lockstep_runner_keeps_the_point_sets_equal_and_counts_disagreements():
    import train_harness as th
    dev = torch.device("cpu")
    cams = th.Cameras(points=600, width=32, height=16, n_cams=2, dev=dev, n_actors=2, actor_points=100)
    for i in range(2):
        cams.gt[i] = torch.full((3, 16, 32), 0.25 + 0.1 * i)
    # thresholds low enough that the stand-in's gradients clone and split some Gaussians
    cfg = dict(densify_grad_threshold_bkgd=1e-7, densify_grad_threshold_obj=1e-7, min_opacity=0.3)
    a, b = _arms(th, cams, cfg)
    out = th.run_training_lockstep(a, b, 12, (5, 10), eval_at=(1, 12))
    ra, rb = out["reference"], out["reference-b"]
    # identical arms: identical decisions, sizes and (bit for bit) losses
    assert ra["losses"] == rb["losses"] and len(ra["losses"]) == 12
    for it in (5, 10):
        assert ra["sizes"][it][0] == rb["sizes"][it][0]
        assert all(s["disagree"] == 0 for s in rb["sizes"][it][1])
    ev = [s for it in (5, 10) for s in ra["sizes"][it][1]]
    assert sum(s["cloned"] + s["split"] for s in ev) > 0 and sum(s["pruned"] for s in ev) > 0
    assert torch.equal(ra["renders"][12], rb["renders"][12])

    # a follower whose statistics differ takes other decisions on its own -- they are counted, the leader's are applied
    a, b = _arms(th, cams, cfg)
    for it in range(1, 5):
        for arm in (a, b):
            arm.iteration(it % 2, cams.gt[it % 2])
    b.stats[0][2].fill_(1e12)  # a huge denominator: the follower's background would clone / split nothing
    a.iteration(0, cams.gt[0], densify_seed=5)
    b.iteration(0, cams.gt[0], densify_seed=5, forced=a.last_masks)
    assert a.sizes() == b.sizes()
    wanted = a.last_densify[0]["cloned"] + a.last_densify[0]["split"]
    assert wanted > 0 and b.last_densify[0]["disagree"] >= wanted
    assert b.last_densify[0]["cloned"] == a.last_densify[0]["cloned"]
    # optimiser state followed the surgery in both arms
    for arm in (a, b):
        for k, p in enumerate(arm.subs):
            for g in arm.opts[k].param_groups:
                assert g["params"][0].shape[0] == p["xyz"].shape[0]
