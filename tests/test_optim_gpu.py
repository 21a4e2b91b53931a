"""GPU parity of the fused Adam step and densification statistics (grpg_adam_step / grpg_densify_stats through
gaussianrpg_b200.optim) with goldens from the reference's own optimiser plumbing, with the numpy oracle, and -- at the
bench size -- with stock torch.optim.Adam / the reference's masked-indexing formulation on the same device."""
import numpy as np
import pytest
import torch

import optim_cases
from gaussianrpg_b200 import optim
from oracle import optim_oracle
from test_optim_oracle_cpu import oracle_lrs

pytestmark = pytest.mark.gpu
NAMES = dict(xyz="xyz", features_dc="f_dc", features_rest="f_rest", opacity="opacity", scaling="scaling",
             rotation="rotation", semantic="semantic")


def _training_setup(params, case):
    """gaussian_model.py:292-304 with FusedAdam in place of torch.optim.Adam."""
    c = optim_cases.OPTIM_CFG
    l = [{'params': [params["xyz"]], 'lr': c["position_lr_init"] * case["scale"], "name": "xyz"},
         {'params': [params["features_dc"]], 'lr': c["feature_lr"], "name": "f_dc"},
         {'params': [params["features_rest"]], 'lr': c["feature_lr"] / 20.0, "name": "f_rest"},
         {'params': [params["opacity"]], 'lr': c["opacity_lr"], "name": "opacity"},
         {'params': [params["scaling"]], 'lr': c["scaling_lr"], "name": "scaling"},
         {'params': [params["rotation"]], 'lr': c["rotation_lr"], "name": "rotation"},
         {'params': [params["semantic"]], 'lr': c["semantic_lr"], "name": "semantic"}]
    return optim.FusedAdam(l, lr=0.0, eps=1e-15)


def _rel(a, b):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return 0.0 if b.size == 0 else float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", list(optim_cases.adam_cases().keys()))
def test_fused_adam_vs_reference_golden(name, cuda_device, golden_dir):
    case = optim_cases.adam_cases()[name]
    gold = np.load(golden_dir / f"optim_adam_{name}.npz")
    params = {k: torch.nn.Parameter(v.to(cuda_device)) for k, v in optim_cases.adam_params(case).items()}
    opt = _training_setup(params, case)
    for s in range(case["steps"]):
        lrs = oracle_lrs(case, case["it0"] + s)
        for g in opt.param_groups:  # update_learning_rate, gaussian_model.py:320-325
            if g["name"] == "xyz":
                g["lr"] = lrs["xyz"]
        for k, g in optim_cases.adam_grads(case, s).items():
            params[k].grad = None if g is None else g.to(cuda_device)
        opt.step()
        opt.zero_grad(set_to_none=True)
    for k in optim_cases.PARAMS:
        st = opt.state[params[k]]
        assert float(st["step"]) == float(gold["step_" + k])
        assert _rel(st["exp_avg"], gold["exp_avg_" + k]) <= 1e-6, k
        assert _rel(st["exp_avg_sq"], gold["exp_avg_sq_" + k]) <= 1e-6, k
        assert _rel(params[k], gold["param_" + k]) <= 1e-6, k
    sd = opt.state_dict()  # checkpoint layout is torch.optim.Adam's
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["name"] == "xyz"


def test_fused_adam_many_optimizers_one_launch_vs_torch(cuda_device):
    """Nine sub-models (1 x 200 k + 8 x 20 k Gaussians) stepped in one launch against stock torch.optim.Adam on
    copies of the same parameters, three iterations; one sub-model has no gradients in the second iteration."""
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    sizes = [200_000] + [20_000] * 8
    ours, theirs, p_ours, p_theirs = [], [], [], []
    for n in sizes:
        case = dict(n=n, M=4, F=5 if n < 100_000 else 1, scale=3.0)
        base = {k: torch.randn(*s, generator=g) for k, s in optim_cases.SHAPES(n, case["M"], case["F"]).items()}
        a = {k: torch.nn.Parameter(v.to(dev)) for k, v in base.items()}
        b = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in base.items()}
        p_ours.append(a); p_theirs.append(b)
        ours.append(_training_setup(a, case))
        theirs.append(torch.optim.Adam([dict(g_, params=[b[k] for k in optim_cases.PARAMS if NAMES[k] == g_["name"]])
                                        for g_ in ({kk: vv for kk, vv in grp.items() if kk in ("lr", "name")}
                                                   for grp in ours[-1].param_groups)], lr=0.0, eps=1e-15))
    for it in range(3):
        for i, (a, b) in enumerate(zip(p_ours, p_theirs)):
            for k in optim_cases.PARAMS:
                if it == 1 and i == 4:
                    a[k].grad = b[k].grad = None
                    continue
                gr = (torch.randn(a[k].shape, generator=g) * 1e-3).to(dev)
                a[k].grad, b[k].grad = gr, gr.clone()
        n_updated = optim.fused_adam_step(ours)
        assert n_updated == (9 if it != 1 else 8) * 6  # the empty `semantic` tensors are not listed
        for o in theirs:
            o.step()
    for a, b, oa, ob in zip(p_ours, p_theirs, ours, theirs):
        for k in optim_cases.PARAMS:
            assert _rel(a[k], b[k].detach().cpu().numpy()) <= 1e-6, k
            if a[k].numel():
                assert _rel(oa.state[a[k]]["exp_avg"], ob.state[b[k]]["exp_avg"].cpu().numpy()) <= 1e-6
                assert _rel(oa.state[a[k]]["exp_avg_sq"], ob.state[b[k]]["exp_avg_sq"].cpu().numpy()) <= 1e-6
                assert float(oa.state[a[k]]["step"]) == float(ob.state[b[k]]["step"])


def test_more_tensors_than_threads_in_one_launch(cuda_device):
    """Descriptor tables longer than the 256-thread block (a scene with background + > 36 actors has > 256 tensors,
    7 per sub-model): 45 sub-models x 7 tensors = 315 rows in ONE launch against stock torch.optim.Adam, and the
    statistics over 300 sub-models against the reference's masked formulation."""
    dev = cuda_device
    g = torch.Generator().manual_seed(9)
    n_sub = 45
    ours, theirs, pa, pb = [], [], [], []
    for i in range(n_sub):
        n = 97 + 131 * i  # ragged: chunks end inside and across the 4096-element tiles
        case = dict(n=n, M=4, F=3, scale=1.0)
        shapes = dict(optim_cases.SHAPES(n, 4, 3))
        shapes["semantic"] = (n, 2)  # non-empty, so that all 7 tensors of a sub-model are listed
        base = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
        a = {k: torch.nn.Parameter(v.to(dev)) for k, v in base.items()}
        b = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in base.items()}
        pa.append(a); pb.append(b)
        ours.append(_training_setup(a, case))
        theirs.append(torch.optim.Adam([dict(lr=grp["lr"], name=grp["name"],
                                             params=[b[k] for k in optim_cases.PARAMS if NAMES[k] == grp["name"]])
                                        for grp in ours[-1].param_groups], lr=0.0, eps=1e-15))
    for it in range(2):
        for a, b in zip(pa, pb):
            for k in optim_cases.PARAMS:
                gr = (torch.randn(a[k].shape, generator=g) * 1e-2).to(dev)
                a[k].grad, b[k].grad = gr, gr.clone()
        assert optim.fused_adam_step(ours) == n_sub * 7 > 256
        for o in theirs:
            o.step()
    for a, b in zip(pa, pb):
        for k in optim_cases.PARAMS:
            assert _rel(a[k], b[k].detach().cpu().numpy()) <= 1e-6, k

    sizes = [5 + (37 * i) % 700 for i in range(300)]
    radii, grad, subs = optim_cases.stats_inputs(sizes, seed=4)
    radii, grad = radii.to(dev), grad.to(dev)
    stats = [optim.DensifyStats(*(s[k].to(dev) for k in ("max_radii2D", "xyz_gradient_accum", "denom"))) for s in subs]
    want = [{k: s[k].to(dev).clone() for k in s} for s in subs]
    optim.update_densification_stats(stats, radii, grad)
    off, vis_all, rf = 0, radii > 0, radii.float()
    for w, n in zip(want, sizes):
        vis, gg = vis_all[off:off + n], grad[off:off + n]
        w["max_radii2D"][vis] = torch.max(w["max_radii2D"][vis], rf[off:off + n][vis])
        w["xyz_gradient_accum"][vis, 0:1] += torch.norm(gg[vis, :2], dim=-1, keepdim=True)
        w["xyz_gradient_accum"][vis, 1:2] += torch.norm(gg[vis, 2:], dim=-1, keepdim=True)
        w["denom"][vis] += 1
        off += n
    for s, w in zip(stats, want):
        assert torch.equal(s.max_radii2D, w["max_radii2D"]) and torch.equal(s.denom, w["denom"])
        assert torch.allclose(s.xyz_gradient_accum, w["xyz_gradient_accum"], rtol=3e-7, atol=0)


def test_densification_surgery_on_fused_adam_state(cuda_device):
    """The reference's optimiser-state surgery (gaussian_model.py:394-470) works on FusedAdam: prune then append."""
    case = dict(n=100, M=4, F=1, scale=1.0)
    params = {k: torch.nn.Parameter(torch.randn(*s, device=cuda_device)) for k, s in optim_cases.SHAPES(100, 4, 1).items()}
    opt = _training_setup(params, case)
    for p in params.values():
        p.grad = torch.randn_like(p)
    opt.step()
    mask = torch.rand(100, device=cuda_device) < 0.5
    for group in opt.param_groups:  # _prune_optimizer, :418-436
        stored = opt.state.get(group['params'][0], None)
        stored["exp_avg"] = stored["exp_avg"][mask]
        stored["exp_avg_sq"] = stored["exp_avg_sq"][mask]
        del opt.state[group['params'][0]]
        group["params"][0] = torch.nn.Parameter(group["params"][0][mask].requires_grad_(True))
        opt.state[group['params'][0]] = stored
    for group in opt.param_groups:
        group["params"][0].grad = torch.randn_like(group["params"][0])
    opt.step()  # runs on the pruned tensors with the carried-over moments and step counts
    assert all(float(opt.state[g["params"][0]]["step"]) == 2.0 for g in opt.param_groups)


@pytest.mark.parametrize("name", list(optim_cases.stats_cases().keys()))
def test_densify_stats_vs_reference_golden(name, cuda_device, golden_dir):
    sizes = optim_cases.stats_cases()[name]
    gold = np.load(golden_dir / f"optim_stats_{name}.npz")
    radii, grad, subs = optim_cases.stats_inputs(sizes)
    stats = [optim.DensifyStats(*(s[k].to(cuda_device) for k in ("max_radii2D", "xyz_gradient_accum", "denom"))) for s in subs]
    optim.update_densification_stats(stats, radii.to(cuda_device), grad.to(cuda_device))
    for k, s in enumerate(stats):
        assert np.array_equal(s.max_radii2D.cpu().numpy(), gold[f"max_radii2D_{k}"])
        assert np.array_equal(s.denom.cpu().numpy(), gold[f"denom_{k}"])
        np.testing.assert_allclose(s.xyz_gradient_accum.cpu().numpy(), gold[f"xyz_gradient_accum_{k}"], rtol=3e-7, atol=0)


def test_densify_stats_bench_size_vs_masked_indexing(cuda_device):
    """2 M rows over 9 sub-models against the reference's masked read-modify-writes on the same device."""
    dev = cuda_device
    sizes = [1_840_000] + [20_000] * 8
    radii, grad, subs = optim_cases.stats_inputs(sizes, seed=3)
    radii, grad = radii.to(dev), grad.to(dev)
    stats = [optim.DensifyStats(*(s[k].to(dev) for k in ("max_radii2D", "xyz_gradient_accum", "denom"))) for s in subs]
    want = [{k: s[k].to(dev).clone() for k in s} for s in subs]
    optim.update_densification_stats(stats, radii, grad)
    off, vis_all, rf = 0, radii > 0, radii.float()
    for w, n in zip(want, sizes):  # street_gaussian_model.py:555-578
        vis, g = vis_all[off:off + n], grad[off:off + n]
        w["max_radii2D"][vis] = torch.max(w["max_radii2D"][vis], rf[off:off + n][vis])
        w["xyz_gradient_accum"][vis, 0:1] += torch.norm(g[vis, :2], dim=-1, keepdim=True)
        w["xyz_gradient_accum"][vis, 1:2] += torch.norm(g[vis, 2:], dim=-1, keepdim=True)
        w["denom"][vis] += 1
        off += n
    for s, w in zip(stats, want):
        assert torch.equal(s.max_radii2D, w["max_radii2D"]) and torch.equal(s.denom, w["denom"])
        assert torch.allclose(s.xyz_gradient_accum, w["xyz_gradient_accum"], rtol=3e-7, atol=0)


def test_argument_errors(cuda_device):
    p = torch.nn.Parameter(torch.randn(4, 3))
    p.grad = torch.randn(4, 3)
    with pytest.raises(RuntimeError):
        optim.FusedAdam([p], lr=1e-3).step()  # CPU parameter: no CPU path
    with pytest.raises(NotImplementedError):
        q = torch.nn.Parameter(torch.randn(4, 3, device=cuda_device)); q.grad = torch.randn_like(q)
        optim.FusedAdam([q], lr=1e-3, amsgrad=True).step()
    with pytest.raises(RuntimeError):
        optim.update_densification_stats([], torch.ones(3, dtype=torch.int32, device=cuda_device),
                                         torch.zeros(3, 3, device=cuda_device))
