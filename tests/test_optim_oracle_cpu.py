"""The rank-3 oracle (numpy restatement of torch's Adam update, the xyz learning-rate schedule and the densification
statistics) against goldens produced by the reference's own GaussianModel / StreetGaussianModel methods on the CPU."""
import numpy as np
import pytest

import optim_cases
from oracle import optim_oracle

GROUP_LR = dict(xyz=None, features_dc="feature_lr", features_rest="feature_lr", opacity="opacity_lr", scaling="scaling_lr",
                rotation="rotation_lr", semantic="semantic_lr")


def oracle_lrs(case, it):
    c = optim_cases.OPTIM_CFG
    lrs = {}
    for k in optim_cases.PARAMS:
        if k == "xyz":  # gaussian_model.py:305-310,320-325
            lrs[k] = optim_oracle.expon_lr(it, c["position_lr_init"] * case["scale"], c["position_lr_final"] * case["scale"],
                                           c["position_lr_delay_mult"], c["position_lr_max_steps"])
        else:
            lrs[k] = c[GROUP_LR[k]] / (20.0 if k == "features_rest" else 1.0)  # :294-300
    return lrs


def run_oracle_adam(case):
    p = {k: v.numpy() for k, v in optim_cases.adam_params(case).items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v2 = {k: np.zeros_like(v) for k, v in p.items()}
    steps = {k: 0 for k in p}
    all_lrs = []
    for s in range(case["steps"]):
        lrs = oracle_lrs(case, case["it0"] + s)
        all_lrs.append([lrs[k] for k in optim_cases.PARAMS])
        for k, g in optim_cases.adam_grads(case, s).items():
            if g is None:
                continue
            steps[k] += 1
            p[k], m[k], v2[k] = optim_oracle.adam_step(p[k], g.numpy(), m[k], v2[k], steps[k], lrs[k])
    return p, m, v2, steps, np.array(all_lrs)


@pytest.mark.parametrize("name", list(optim_cases.adam_cases().keys()))
def test_adam_oracle_matches_reference_optimizer(name, golden_dir):
    case = optim_cases.adam_cases()[name]
    gold = np.load(golden_dir / f"optim_adam_{name}.npz")
    assert float(gold["eps"]) == 1e-15 and tuple(gold["betas"]) == (0.9, 0.999)  # gaussian_model.py:304
    p, m, v, steps, lrs = run_oracle_adam(case)
    assert np.allclose(lrs, gold["lrs"], rtol=1e-12, atol=0)
    for k in optim_cases.PARAMS:
        assert steps[k] == int(gold["step_" + k])
        # float32 operation order is restated exactly up to fused-multiply-add contraction: 1-2 ulp of the operands,
        # i.e. a few 1e-7 of the tensor's scale (more, relatively, on elements where g - m cancels)
        for a, b in ((m[k], gold["exp_avg_" + k]), (v[k], gold["exp_avg_sq_" + k]), (p[k], gold["param_" + k])):
            if b.size:
                assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
                np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-6 * np.abs(b).max())


@pytest.mark.parametrize("name", list(optim_cases.stats_cases().keys()))
def test_stats_oracle_matches_reference_methods(name, golden_dir):
    sizes = optim_cases.stats_cases()[name]
    gold = np.load(golden_dir / f"optim_stats_{name}.npz")
    radii, grad, subs = optim_cases.stats_inputs(sizes)
    out = optim_oracle.densify_stats(radii.numpy(), grad.numpy(), [{k: v.numpy() for k, v in s.items()} for s in subs])
    for k, o in enumerate(out):
        assert np.array_equal(o["max_radii2D"], gold[f"max_radii2D_{k}"])
        assert np.array_equal(o["denom"], gold[f"denom_{k}"])
        np.testing.assert_allclose(o["xyz_gradient_accum"], gold[f"xyz_gradient_accum_{k}"], rtol=3e-7, atol=0)
