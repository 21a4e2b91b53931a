"""BASELINE config #4 as a real training run (tools/train_harness.py): compose -> rasterizer -> L1 + 0.2 DSSIM ->
backward -> densification statistics -> densify / clone / split / prune -> Adam, several hundred iterations, with this
repository's operators in one arm and the reference's formulation of every stage (stock PyTorch + the UNMODIFIED
reference rasterizer from oracle/_ref) in the other.  Same initial parameters, same ground truth, same iteration count,
densification at the same iterations: the loss trajectories, the Gaussian counts after every densify and the rendered
frames must agree."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu


def test_training_trajectory_matches_the_reference_arm(cuda_device):
    import train_harness as th
    import diff_gaussian_rasterization as ours_dgr
    ref_dgr = th.load_reference_extension()
    if ref_dgr is None:
        pytest.skip("oracle/_ref (reference extension) not built")
    dev = cuda_device
    iters, densify_at, eval_at = 320, (100, 200, 300), (1, 99, 200, 320)
    # lower gradient thresholds than the street config so that a small scene clones, splits AND prunes within 300
    # iterations (the bench-size run, tools/train_config4.py, uses the YAML values as they are)
    cfg = dict(densify_grad_threshold_bkgd=0.00015, densify_grad_threshold_obj=0.00005)
    cams = th.Cameras(points=60_000, width=480, height=320, n_cams=4, dev=dev, n_actors=4, actor_points=3000)
    th.make_ground_truth(cams, ours_dgr, dev)
    runs = {}
    for name, dgr in (("ours", ours_dgr), ("reference", ref_dgr)):
        arm = th.Arm(name, dgr, cams, dev, cfg=cfg, bkgd_extent=3.0, actor_extent=2.0)
        p0 = arm.total()
        runs[name] = th.run_training(arm, iters, densify_at, eval_at=eval_at)
        runs[name]["P_end"], runs[name]["P_start"] = arm.total(), p0
    o, r = runs["ours"], runs["reference"]
    # the run really densified: P changes at every event, in both directions over the run
    events = [o["sizes"][it][1] for it in densify_at]
    assert sum(s["cloned"] for e in events for s in e) > 0
    assert sum(s["split"] for e in events for s in e) > 0
    assert sum(s["pruned"] for e in events for s in e) > 0
    assert o["P_end"] != o["P_start"]
    # Gaussian counts after each densify: equal up to threshold flips of borderline Gaussians (the accumulated
    # statistics differ by ~1e-5 relative between the arms).  The first event sees the same model in both arms: 0.1 %;
    # afterwards the arms hold slightly different point sets and the flips compound: 3 %.  The counts are printed.
    for n_ev, it in enumerate(densify_at):
        so, sr = o["sizes"][it][0], r["sizes"][it][0]
        for a, b in zip(so, sr):
            assert abs(a - b) <= max(3, int((0.001 if n_ev == 0 else 0.03) * b)), (it, so, sr)
    print("P after densify (ours / reference):", {it: (sum(o["sizes"][it][0]), sum(r["sizes"][it][0])) for it in densify_at})
    # Loss trajectories.  Up to the first densify both arms hold the same model: the curves agree iteration by
    # iteration.  One borderline Gaussian flipping across a threshold (65944 vs 65945 points after the first event in the
    # measured run) shifts the split's random draws, so afterwards the arms train different -- equally good -- point
    # sets: there the curves are compared as 20-iteration means, and both must go down.
    lo, lr = torch.tensor(o["losses"]), torch.tensor(r["losses"])
    first = densify_at[0] - 1
    assert float((lo[:first] - lr[:first]).abs().max()) <= 5e-3 * float(lr[0])  # float noise on a loss that falls 25x
    wo, wr = lo[first:first + 220].reshape(-1, 20).mean(1), lr[first:first + 220].reshape(-1, 20).mean(1)
    assert float(((wo - wr).abs() / wr).max()) <= 0.2, ((wo - wr).abs() / wr)
    # training works in both arms (a densify event itself makes the loss jump: compare the stretch before the first one)
    assert lo[first - 20:first].mean() < 0.8 * lo[:20].mean() and lr[first - 20:first].mean() < 0.8 * lr[:20].mean()
    # rendered frames of a fixed camera at the checkpoints: PSNR(ours, reference)
    report = {}
    for it in eval_at:
        report[it] = th.psnr(o["renders"][it], r["renders"][it])
    print("PSNR(ours, reference) of the evaluation camera:", report)
    assert report[1] >= 80.0 and report[99] >= 60.0  # same model in both arms up to the first densify: float noise only
    assert min(report.values()) >= 25.0  # afterwards the arms hold different point sets of the same scene
    # both arms end equally close to the ground truth
    po = th.psnr(o["renders"][iters].clamp(0, 1), cams.gt[0])
    pr = th.psnr(r["renders"][iters].clamp(0, 1), cams.gt[0])
    print("PSNR vs ground truth at the end (ours, reference):", po, pr)
    assert abs(po - pr) <= 1.0
