"""The compose oracle (numpy float64, hand-derived backward) against the goldens produced by the reference's own
StreetGaussianModel getters + autograd (tests/golden/make_compose_golden.py)."""
import numpy as np
import pytest

import compose_cases
from oracle import compose_oracle

PARAMS = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")


def run_oracle(case):
    idfts = [compose_cases.idft_base(t, case["F"]) for t in case["times"]]
    np_sub = lambda s: None if s is None else {k: v.numpy() for k, v in s.items()}
    args = (np_sub(case["bkgd"]), [np_sub(a) for a in case["actors"]], case["obj_rots"].numpy(), case["obj_trans"].numpy(),
            idfts, [f.numpy() for f in case["flips"]])
    out = compose_oracle.compose(*args)
    w = {k: v.numpy() for k, v in compose_cases.out_weights(case).items()}
    return out, compose_oracle.compose_bwd(*args, w)


def close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err <= tol, err


@pytest.mark.parametrize("name", list(compose_cases.cases().keys()))
def test_oracle_matches_reference_getters(name, golden_dir):
    case = compose_cases.cases()[name]
    gold = np.load(golden_dir / f"compose_{name}.npz")
    assert np.allclose(np.abs(gold["flip_quat"].ravel()), np.abs(compose_oracle.FLIP_QUAT))
    out, (grads, g_rots, g_trans) = run_oracle(case)
    for k, v in out.items():
        close(v, gold["out_" + k], 2e-6)  # the goldens are float32
    tags = (["bkgd"] if case["bkgd"] is not None else []) + [f"actor{k}" for k in range(len(case["actors"]))]
    for tag, d in zip(tags, grads):
        for p in PARAMS:
            close(d[p], gold[f"grad_{tag}_{p}"], 2e-5)
    if case["actors"]:
        close(g_rots, gold["grad_obj_rots"], 1e-4)   # float32 sums over the actor's Gaussians in the golden
        close(g_trans, gold["grad_obj_trans"], 1e-4)


def test_idft_base_matches_reference_lines():
    b = compose_cases.idft_base(0.37, 5)
    t = 0.37
    want = [1.0, np.sin(np.pi * t * 2), np.cos(np.pi * t * 2), np.sin(np.pi * t * 4), np.cos(np.pi * t * 4)]
    assert np.allclose(b, want, atol=1e-6)


def test_street_scene_graph_composes_back_to_the_scene():
    """gaussianrpg_b200.synthetic.street_scene_graph (the sub-model form of BASELINE config #3 used by
    tools/train_iter_bench.py) must compose, by the reference's rules, into exactly the rasterizer inputs of
    synthetic.street_scene."""
    import torch
    from gaussianrpg_b200 import synthetic
    sc = synthetic.street_scene(P=120_000, n_actors=8, actor_points=2_000)
    bk, actors, rots, trans, idft = synthetic.street_scene_graph(sc)
    assert bk["xyz"].shape[0] == sc.extra["n_bkgd"] and len(actors) == 8 and actors[0]["features_dc"].shape[1] == 5
    np_ = lambda d: {k: v.numpy() for k, v in d.items()}  # noqa: E731
    out = compose_oracle.compose(np_(bk), [np_(a) for a in actors], rots.numpy(), trans.numpy(), idft,
                                 [np.zeros(a["xyz"].shape[0], dtype=bool) for a in actors])
    for k, ref in (("xyz", sc.means3D), ("rotation", sc.rotations), ("scaling", sc.scales), ("opacity", sc.opacities),
                   ("features", sc.shs)):
        assert np.abs(out[k] - ref.numpy()).max() <= 5e-6 * max(1.0, float(ref.abs().max())), k
