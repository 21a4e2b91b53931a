"""BASELINE config #4 as a real training run (tools/train_harness.py): compose -> rasterizer -> L1 + 0.2 DSSIM ->
backward -> densification statistics -> densify / clone / split / prune -> Adam, several hundred iterations, with this
repository's operators in one arm and the reference's formulation of every stage (stock PyTorch + the UNMODIFIED
reference rasterizer from oracle/_ref) in the other.  Same initial parameters, same ground truth, same iteration count,
densification at the same iterations.  The arms run in LOCK STEP (train_harness.run_training_lockstep): at every densify
event each arm takes its own clone / split / prune decisions, the test counts the Gaussians on which they differ (a
borderline statistic on either side of a threshold) and the reference arm's decisions are applied in both, so the arms
hold the same point set throughout and loss, rendered frame and final quality can be compared iteration by iteration
over the whole run.  (Free-running arms drift apart after the first borderline Gaussian -- 65938 vs 65942 points after
the first event in the measured run -- because the split's random draws shift; that comparison, at bench size, is
tools/train_config4.py.)"""
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
    ref_arm = th.Arm("reference", ref_dgr, cams, dev, cfg=cfg, bkgd_extent=3.0, actor_extent=2.0)
    our_arm = th.Arm("ours", ours_dgr, cams, dev, cfg=cfg, bkgd_extent=3.0, actor_extent=2.0)
    p_start = our_arm.total()
    runs = th.run_training_lockstep(ref_arm, our_arm, iters, densify_at, eval_at=eval_at)
    o, r = runs["ours"], runs["reference"]
    # the run really densified: P changes at every event, in both directions over the run
    events = [o["sizes"][it][1] for it in densify_at]
    assert sum(s["cloned"] for e in events for s in e) > 0
    assert sum(s["split"] for e in events for s in e) > 0
    assert sum(s["pruned"] for e in events for s in e) > 0
    assert our_arm.total() != p_start and our_arm.sizes() == ref_arm.sizes()
    # Decisions: the Gaussians this repository's arm would have cloned / split / pruned differently from the reference
    # arm, per event.  The accumulated statistics of the arms differ by float noise, so only Gaussians sitting on a
    # threshold can flip: at most 0.1 % of the model at the first event (same model, ~1e-5 relative noise), 1 % later
    # (the parameters have been through hundreds of Adam steps, which amplify noise on near-zero gradients).
    for n_ev, it in enumerate(densify_at):
        sizes, rep = o["sizes"][it]
        disagree, total = sum(s["disagree"] for s in rep), sum(s["before"] for s in rep)
        print(f"densify at {it}: P {total} -> {sum(sizes)}, decisions differing from the reference arm: {disagree}")
        assert disagree <= max(3, int((0.001 if n_ev == 0 else 0.01) * total)), (it, disagree, total)
        assert sizes == r["sizes"][it][0]
    # Loss trajectories, whole run, iteration by iteration and as 20-iteration means
    lo, lr = torch.tensor(o["losses"]), torch.tensor(r["losses"])
    first = densify_at[0] - 1
    assert float((lo[:first] - lr[:first]).abs().max()) <= 5e-3 * float(lr[0])  # float noise on a loss that falls 25x
    rel = ((lo - lr).abs() / lr)
    wo, wr = lo.reshape(-1, 20).mean(1), lr.reshape(-1, 20).mean(1)
    print("loss: max relative difference per iteration %.3e, per 20-iteration mean %.3e" % (float(rel.max()), float(((wo - wr).abs() / wr).max())))
    assert float(rel.max()) <= 0.1 and float(((wo - wr).abs() / wr).max()) <= 0.05
    # training works in both arms (a densify event itself makes the loss jump: compare the stretch before the first one)
    assert lo[first - 20:first].mean() < 0.8 * lo[:20].mean() and lr[first - 20:first].mean() < 0.8 * lr[:20].mean()
    # rendered frames of a fixed camera at the checkpoints: PSNR(ours, reference)
    report = {it: th.psnr(o["renders"][it], r["renders"][it]) for it in eval_at}
    print("PSNR(ours, reference) of the evaluation camera:", report)
    assert report[1] >= 80.0 and report[99] >= 60.0  # float noise only
    assert min(report.values()) >= 40.0  # same point set through every densify event (200 is right after one)
    # both arms end equally close to the ground truth
    po = th.psnr(o["renders"][iters].clamp(0, 1), cams.gt[0])
    pr = th.psnr(r["renders"][iters].clamp(0, 1), cams.gt[0])
    print("PSNR vs ground truth at the end (ours, reference):", po, pr)
    assert abs(po - pr) <= 0.5
