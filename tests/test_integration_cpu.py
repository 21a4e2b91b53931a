"""Plumbing of gaussianrpg_b200.integration.install on the reference's OWN StreetGaussianModel class (imported from
/root/reference on the CPU, constructors bypassed as in tests/golden/make_compose_golden.py), with CPU stand-ins for
the CUDA operators: the patched getters / optimiser step / statistics must reproduce the original methods.  Skipped
where the reference tree is absent (the GPU box); the operators themselves are covered by the -m gpu tests."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not Path("/root/reference/lib/models/street_gaussian_model.py").exists(),
                                reason="reference tree not present")
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def _torch_compose(bk, actors, rots, trans, idft, flips):
    from test_compose_gpu import _torch_reference
    return _torch_reference(bk, actors, rots, trans, idft, flips)


@pytest.fixture()
def patched():
    import _ref_import
    import compose_cases
    import make_compose_golden
    from gaussianrpg_b200 import integration
    sgm = _ref_import.load("lib.models.street_gaussian_model")
    case = compose_cases.cases()["street_small"]
    with _ref_import.cpu_device():
        model, obj_rots, obj_trans = make_compose_golden.build_reference_model(case)
        want = dict(xyz=model.get_xyz, rotation=model.get_rotation, scaling=model.get_scaling, opacity=model.get_opacity,
                    features=model.get_features)
    calls = {"adam": 0, "stats": 0}

    def adam_step(opts):
        calls["adam"] += 1
        for o in opts:
            o.step()

    def stats_update(stats, radii, grad, visibility_filter=None, what=3):
        """CPU stand-in for optim.update_densification_stats (same signature and semantics)."""
        calls["stats"] += 1
        off = 0
        for s in stats:
            n = s.denom.shape[0]
            vis = (visibility_filter if visibility_filter is not None else radii > 0)[off:off + n].bool()
            if what & 1:
                s.max_radii2D[vis] = torch.max(s.max_radii2D[vis], radii[off:off + n].float()[vis])
            if what & 2:
                g = grad[off:off + n]
                s.xyz_gradient_accum[vis, 0:1] += torch.norm(g[vis, :2], dim=-1, keepdim=True)
                s.xyz_gradient_accum[vis, 1:2] += torch.norm(g[vis, 2:], dim=-1, keepdim=True)
                s.denom[vis] += 1
            off += n

    orig = integration.install(sgm.StreetGaussianModel, compose=_torch_compose, idft_base=compose_cases.idft_base,
                               adam_step=adam_step, stats_update=stats_update)
    model._grpg_orig = orig
    try:
        yield sgm, model, case, want, obj_rots, obj_trans, calls
    finally:
        integration.uninstall(sgm.StreetGaussianModel, orig)


def test_patched_getters_reproduce_the_reference_getters(patched):
    import _ref_import
    sgm, model, case, want, obj_rots, obj_trans, calls = patched
    model._grpg_composed = None  # what the patched parse_camera leaves behind
    with _ref_import.cpu_device():
        got = dict(xyz=model.get_xyz, rotation=model.get_rotation, scaling=model.get_scaling, opacity=model.get_opacity,
                   features=model.get_features)
    for k in want:
        assert got[k].shape == want[k].shape
        assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-6), k
    assert model.get_xyz is got["xyz"]  # one compose per parse_camera, cached
    # autograd reaches the per-actor pose leaves through the row slicing of the expanded pose
    (got["xyz"].sum() + got["rotation"].sum()).backward()
    assert obj_rots.grad is not None and obj_rots.grad.abs().sum() > 0
    assert obj_trans.grad is not None and torch.allclose(obj_trans.grad, torch.tensor(
        [[float(a["xyz"].shape[0])] * 3 for a in case["actors"]]))


def test_fallback_to_the_original_getters(patched):
    import _ref_import
    sgm, model, case, want, *_ = patched
    model._grpg_composed = None
    model.use_pose_correction = True  # not covered by the fused path -> original code (would need pose_correction)
    model.pose_correction = type("PC", (), {"correct_gaussian_xyz": staticmethod(lambda cam, x: x),
                                            "correct_gaussian_rotation": staticmethod(lambda cam, r: r)})()
    model.viewpoint_camera = None
    with _ref_import.cpu_device():
        assert torch.equal(model.get_xyz, want["xyz"])
    assert getattr(model, "_grpg_composed", None) is None


def test_patched_optimizer_step_and_statistics(patched):
    import _ref_import
    import optim_cases
    sgm, model, case, want, obj_rots, obj_trans, calls = patched
    gm = _ref_import.load("lib.models.gaussian_model")
    gm.cfg.optim = optim_cases.optim_namespace()
    names = (["background"] if case["bkgd"] is not None else []) + list(model.graph_obj_list)
    model.model_name_id = {n: i for i, n in enumerate(names)}
    model.actor_pose = model.sky_cubemap = model.color_correction = model.pose_correction = None
    sizes = []
    with _ref_import.cpu_device():
        for n in names:
            m = getattr(model, n)
            m._semantic = torch.nn.Parameter(torch.zeros(m._xyz.shape[0], 0))
            m.spatial_lr_scale, m.max_sh_degree = 2.0, 1
            m.training_setup()
            m.max_radii2D = torch.zeros(m._xyz.shape[0])
            sizes.append(m._xyz.shape[0])
            for p in (m._xyz, m._opacity):
                p.grad = torch.ones_like(p) * 1e-3
    before = [getattr(model, n)._xyz.detach().clone() for n in names]
    model.update_optimizer()
    assert calls["adam"] == 1  # one call for all sub-models
    for n, b in zip(names, before):
        m = getattr(model, n)
        assert not torch.equal(m._xyz.detach(), b) and m._xyz.grad is None  # stepped, zero_grad(set_to_none=True)
    # statistics: patched pair of calls == the original pair
    radii, grad, _ = optim_cases.stats_inputs(sizes, seed=8)
    model.graph_gaussian_range, idx = {}, 0
    for n, sz in zip(names, sizes):
        model.graph_gaussian_range[n] = [idx, idx + sz - 1]
        idx += sz
    vp = torch.zeros(sum(sizes), 3, requires_grad=True)
    vp.grad = grad.clone()
    model.set_max_radii2D(radii, radii > 0)
    # like the reference, each call takes effect immediately (set_max_radii2D alone must already update max_radii2D)
    assert any(float(getattr(model, n).max_radii2D.max()) > 0 for n in names)
    model.add_densification_stats(vp, radii > 0)
    assert calls["stats"] == 2  # one launch per call, over all sub-models
    got = [(getattr(model, n).max_radii2D.clone(), getattr(model, n).xyz_gradient_accum.clone(), getattr(model, n).denom.clone())
           for n in names]
    # expected: the oracle, which tests/test_optim_oracle_cpu.py pins to the reference's own two methods
    from oracle import optim_oracle
    want_stats = optim_oracle.densify_stats(radii.numpy(), grad.numpy(), [dict(max_radii2D=np.zeros(s, np.float32),
                                            xyz_gradient_accum=np.zeros((s, 2), np.float32), denom=np.zeros((s, 1), np.float32))
                                            for s in sizes])
    for (mr, acc, den), w in zip(got, want_stats):
        assert np.array_equal(mr.numpy(), w["max_radii2D"]) and np.array_equal(den.numpy(), w["denom"])
        assert np.allclose(acc.numpy(), w["xyz_gradient_accum"], rtol=1e-6)
    # a caller's own mask (not radii > 0) is honoured: patched pair == the reference's original pair on a copy
    mask = (radii > 0) & (torch.arange(radii.shape[0]) % 3 != 0)
    snap = [(m.max_radii2D.clone(), m.xyz_gradient_accum.clone(), m.denom.clone()) for m in (getattr(model, n) for n in names)]
    model.set_max_radii2D(radii * 2, mask)
    model.add_densification_stats(vp, mask)
    patched_out = [(m.max_radii2D.clone(), m.xyz_gradient_accum.clone(), m.denom.clone()) for m in (getattr(model, n) for n in names)]
    for n, (a, b, c) in zip(names, snap):
        m = getattr(model, n)
        m.max_radii2D, m.xyz_gradient_accum, m.denom = a, b, c
    model._grpg_orig["set_max_radii2D"](model, radii * 2, mask)
    model._grpg_orig["add_densification_stats"](model, vp, mask)
    for n, (a, b, c) in zip(names, patched_out):
        m = getattr(model, n)
        assert torch.equal(m.max_radii2D, a) and torch.equal(m.denom, c) and torch.allclose(m.xyz_gradient_accum, b)
