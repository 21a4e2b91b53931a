"""Golden vectors for the fused Adam step and the densification statistics, produced by the
reference's OWN code on the CPU:

    python tests/golden/make_optim_golden.py          # needs /root/reference; writes tests/golden/optim_*.npz

* Adam: a reference `GaussianModel` (built without its constructor) runs its real `training_setup()`
  (lib/models/gaussian_model.py:286-318: the seven groups, their learning rates, eps = 1e-15, the xyz schedule),
  then `update_learning_rate(it)` + `update_optimizer()` for a few iterations with seeded gradients.
* statistics: a reference `StreetGaussianModel` runs its real `set_max_radii2D` / `add_densification_stats`
  (street_gaussian_model.py:555-578) over `graph_gaussian_range`.
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import optim_cases  # noqa: E402
import _ref_import  # noqa: E402


def make_adam(name, case, out_dir):
    gm = _ref_import.load("lib.models.gaussian_model")
    gm.cfg.optim = optim_cases.optim_namespace()
    model = object.__new__(gm.GaussianModel)
    nn.Module.__init__(model)
    p0 = optim_cases.adam_params(case)
    for k, v in p0.items():
        setattr(model, "_" + k, nn.Parameter(v.clone()))
    model.setup_functions()
    model.spatial_lr_scale = case["scale"]
    model.max_sh_degree = 1
    with _ref_import.cpu_device():
        model.training_setup()
    d = {}
    lrs = []
    for step in range(case["steps"]):
        it = case["it0"] + step
        model.update_learning_rate(it)
        lrs.append([g["lr"] for g in model.optimizer.param_groups])
        for k, g in optim_cases.adam_grads(case, step).items():
            getattr(model, "_" + k).grad = None if g is None else g.clone()
        model.update_optimizer()
    d["lrs"] = np.array(lrs, dtype=np.float64)
    d["group_names"] = np.array([g["name"] for g in model.optimizer.param_groups])
    d["eps"] = np.float64(model.optimizer.param_groups[0]["eps"])
    d["betas"] = np.array(model.optimizer.param_groups[0]["betas"], dtype=np.float64)
    for k in optim_cases.PARAMS:
        p = getattr(model, "_" + k)
        st = model.optimizer.state[p]
        d["param_" + k] = p.detach().numpy()
        d["exp_avg_" + k] = st["exp_avg"].numpy()
        d["exp_avg_sq_" + k] = st["exp_avg_sq"].numpy()
        d["step_" + k] = np.float64(float(st["step"]))
    np.savez_compressed(out_dir / f"optim_adam_{name}.npz", **d)
    print("adam", name, d["group_names"], d["lrs"][0], {k: float(d["step_" + k]) for k in optim_cases.PARAMS})


def make_stats(name, sizes, out_dir):
    sgm = _ref_import.load("lib.models.street_gaussian_model")
    radii, grad, subs = optim_cases.stats_inputs(sizes)
    model = object.__new__(sgm.StreetGaussianModel)
    nn.Module.__init__(model)
    model.graph_gaussian_range = {}
    idx = 0
    for k, (n, s) in enumerate(zip(sizes, subs)):
        sub = nn.Module()
        sub.max_radii2D, sub.xyz_gradient_accum, sub.denom = (s[x].clone() for x in ("max_radii2D", "xyz_gradient_accum", "denom"))
        mname = "background" if k == 0 else f"obj_{k:03d}"
        setattr(model, mname, sub)
        model.graph_gaussian_range[mname] = [idx, idx + n - 1]  # parse_camera, :248-262
        idx += n
    vis = radii > 0  # street_gaussian_renderer.py:268
    vp = torch.zeros(sum(sizes), 3, requires_grad=True)
    vp.grad = grad.clone()
    model.set_max_radii2D(radii, vis)
    model.add_densification_stats(vp, vis)
    d = {}
    for k, mname in enumerate(model.graph_gaussian_range):
        sub = getattr(model, mname)
        d[f"max_radii2D_{k}"], d[f"xyz_gradient_accum_{k}"], d[f"denom_{k}"] = (sub.max_radii2D.numpy(), sub.xyz_gradient_accum.numpy(), sub.denom.numpy())
    np.savez_compressed(out_dir / f"optim_stats_{name}.npz", **d)
    print("stats", name, sizes, int(vis.sum()))


def main():
    out_dir = ROOT / "tests" / "golden"
    for name, case in optim_cases.adam_cases().items():
        make_adam(name, case, out_dir)
    for name, sizes in optim_cases.stats_cases().items():
        make_stats(name, sizes, out_dir)


if __name__ == "__main__":
    main()
