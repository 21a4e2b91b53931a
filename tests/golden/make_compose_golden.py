"""Golden vectors for the scene-graph compose op, produced by running the reference's OWN
`StreetGaussianModel.get_xyz / get_rotation / get_scaling / get_opacity / get_features`
(lib/models/street_gaussian_model.py:295-384,438-453) and the sub-model getters they call
(gaussian_model.py:224-251, gaussian_model_bkgd.py:44-67, gaussian_model_actor.py:73-82) on the CPU:

    python tests/golden/make_compose_golden.py        # needs /root/reference; writes tests/golden/compose_*.npz

The model objects are built without their constructors (which need a dataset, a config file and a GPU); every
attribute the getters read is set by hand from tests/compose_cases.py, the getters themselves are untouched.
Gradients come from autograd through those getters for the loss sum(w * out) (compose_cases.out_weights).
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import compose_cases  # noqa: E402
import _ref_import  # noqa: E402

PARAMS = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")


def _bare(cls, sub, **attrs):
    m = object.__new__(cls)
    nn.Module.__init__(m)
    for k in PARAMS:
        setattr(m, "_" + k, nn.Parameter(sub[k].clone()))
    m.setup_functions()
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def build_reference_model(case):
    sgm = _ref_import.load("lib.models.street_gaussian_model")
    gu = _ref_import.load("lib.utils.general_utils")
    model = object.__new__(sgm.StreetGaussianModel)
    nn.Module.__init__(model)
    model.include_background = case["bkgd"] is not None
    model.include_obj = True
    model.include_sky = False
    model.use_pose_correction = False
    model.include_list = (["background"] if case["bkgd"] is not None else [])
    if case["bkgd"] is not None:
        model.background = _bare(sgm.GaussianModelBkgd, case["bkgd"], background_mask=None)
    model.graph_obj_list = []
    model.frame = 7
    for k, sub in enumerate(case["actors"]):
        name = f"obj_{k:03d}"
        # get_features_fourier: time = fourier_scale * (frame - start_frame) / (end_frame - start_frame)
        actor = _bare(sgm.GaussianModelActor, sub, fourier_dim=case["F"], fourier_scale=1.0, start_frame=7,
                      end_frame=8, deformable=False)
        t = case["times"][k]
        actor.start_frame, actor.end_frame, actor.fourier_scale = 7.0 - t, 8.0 - t, 1.0  # (7 - (7 - t)) / 1 = t
        setattr(model, name, actor)
        model.graph_obj_list.append(name)
        model.include_list.append(name)
    model.flip_axis = compose_cases.FLIP_AXIS
    fm = torch.eye(3).float() * -1  # street_gaussian_model.py:59-61
    fm[model.flip_axis, model.flip_axis] = 1
    model.flip_matrix = gu.matrix_to_quaternion(fm.unsqueeze(0))
    # parse_camera (:265-293) expands the per-actor pose to one row per Gaussian and concatenates
    obj_rots = case["obj_rots"].clone().requires_grad_(True)
    obj_trans = case["obj_trans"].clone().requires_grad_(True)
    if case["actors"]:
        model.obj_rots = torch.cat([obj_rots[k].expand(a["xyz"].shape[0], -1) for k, a in enumerate(case["actors"])], 0)
        model.obj_trans = torch.cat([obj_trans[k].unsqueeze(0).expand(a["xyz"].shape[0], -1)
                                     for k, a in enumerate(case["actors"])], 0)
        model.flip_mask = torch.cat(case["flips"], 0)
    return model, obj_rots, obj_trans


def main():
    out_dir = ROOT / "tests" / "golden"
    for name, case in compose_cases.cases().items():
        with _ref_import.cpu_device():
            model, obj_rots, obj_trans = build_reference_model(case)
            outs = dict(xyz=model.get_xyz, rotation=model.get_rotation, scaling=model.get_scaling,
                        opacity=model.get_opacity, features=model.get_features)
            w = compose_cases.out_weights(case)
            loss = sum((outs[k] * w[k]).sum() for k in outs)
            loss.backward()
        d = {"out_" + k: v.detach().numpy() for k, v in outs.items()}
        d["flip_quat"] = model.flip_matrix.numpy()
        if case["actors"]:
            d["grad_obj_rots"], d["grad_obj_trans"] = obj_rots.grad.numpy(), obj_trans.grad.numpy()
        subs = ([("bkgd", model.background)] if case["bkgd"] is not None else []) + \
               [(f"actor{k}", getattr(model, n)) for k, n in enumerate(model.graph_obj_list)]
        for tag, m in subs:
            for p in PARAMS:
                d[f"grad_{tag}_{p}"] = getattr(m, "_" + p).grad.numpy()
        np.savez_compressed(out_dir / f"compose_{name}.npz", **d)
        print(name, compose_cases.total(case), {k: v.shape for k, v in d.items() if k.startswith("out_")},
              "flip_quat", d["flip_quat"].ravel())


if __name__ == "__main__":
    main()
