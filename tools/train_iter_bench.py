"""BASELINE config #4 harness: one TRAINING ITERATION of the reference's loop (train.py:64-307) around the op, on the
2 M-Gaussian street scene stored as sub-models (1 background + 8 actors), both arms in one process:

  compose (street_gaussian_model.py:295-384)  ->  rasterizer forward  ->  L1 + 0.2 DSSIM (train.py:116-118)
  ->  backward  ->  densification statistics (train.py:277-281)  ->  per-sub-model Adam (train.py:305-307)

ours      : compose_scene + GaussianRasterizer + l1_ssim_loss + update_densification_stats + fused_adam_step
reference : the reference's formulation of every stage -- stock PyTorch ops for compose / loss / statistics, nine
            torch.optim.Adam(eps=1e-15), and the UNMODIFIED reference rasterizer extension (oracle/_ref).
train.py itself cannot run in this image (plyfile, bidict, nvdiffrast, simple_knn are absent); densify/prune (every 100
iterations, plain tensor surgery in both arms) is not part of the timed iteration.  Also reports the gradient match of
the first iteration.  Prints one JSON line.

    python tools/train_iter_bench.py [--iters 30] [--points 2000000]
"""
import argparse
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import optim_cases
from gaussianrpg_b200 import synthetic, scene_compose, loss_utils, optim
from test_compose_gpu import _torch_reference as torch_compose
from test_loss_gpu import _torch_reference as torch_loss
from test_optim_gpu import _training_setup, NAMES

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--points", type=int, default=2_000_000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1280)
args = ap.parse_args()
dev = torch.device("cuda:0")
sc = synthetic.street_scene(P=args.points, W=args.width, H=args.height)
bk, actors, rots0, trans0, idft = synthetic.street_scene_graph(sc)
P = sc.means3D.shape[0]
sizes = [bk["xyz"].shape[0]] + [a["xyz"].shape[0] for a in actors]
ORDER = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")


def load_ref():
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_ref", ROOT / "oracle" / "build_ref.py")
    build_ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build_ref)
    return build_ref.load() if build_ref.available() else None


class Arm:
    def __init__(self, name, dgr):
        self.name, self.dgr = name, dgr
        self.subs = []
        for d in [bk] + actors:
            p = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in d.items()}
            p["semantic"] = torch.nn.Parameter(torch.zeros(d["xyz"].shape[0], 0, device=dev))
            self.subs.append(p)
        self.rots = rots0.to(dev).requires_grad_(True)   # actor_pose parameters stay on their own optimiser in the
        self.trans = trans0.to(dev).requires_grad_(True)  # reference (actor_pose.py); not stepped here
        self.flips = [torch.zeros(n, dtype=torch.bool, device=dev) for n in sizes[1:]]
        cases = [dict(n=n, M=4, F=d["features_dc"].shape[1], scale=3.0) for n, d in zip(sizes, [bk] + actors)]
        if name == "ours":
            self.opts = [_training_setup(p, c) for p, c in zip(self.subs, cases)]
        else:
            proto = _training_setup({k: torch.nn.Parameter(torch.zeros(1)) for k in optim_cases.PARAMS}, cases[0])
            self.opts = [torch.optim.Adam([dict(lr=g["lr"] if g["name"] != "xyz" else 0.00016 * 3.0, name=g["name"],
                                               params=[p[k] for k in optim_cases.PARAMS if NAMES[k] == g["name"]])
                                          for g in proto.param_groups], lr=0.0, eps=1e-15) for p in self.subs]
        self.stats = [optim.DensifyStats(torch.zeros(n, device=dev), torch.zeros(n, 2, device=dev), torch.zeros(n, 1, device=dev))
                      for n in sizes]
        self.settings = dgr.GaussianRasterizationSettings(
            image_height=sc.height, image_width=sc.width, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=sc.bg.to(dev),
            scale_modifier=1.0, viewmatrix=sc.viewmatrix.to(dev), projmatrix=sc.projmatrix.to(dev), sh_degree=sc.sh_degree,
            campos=sc.campos.to(dev), prefiltered=False, debug=False)

    def compose(self):
        subm = [scene_compose.SubModel(*(p[k] for k in ORDER)) for p in self.subs]
        if self.name == "ours":
            return scene_compose.compose_scene(subm[0], subm[1:], self.rots, self.trans, idft, None)
        return torch_compose(subm[0], subm[1:], self.rots, self.trans, idft, self.flips)

    def render(self):
        out = self.compose()
        means2D = torch.zeros(P, 3, device=dev, requires_grad=True)  # street_gaussian_renderer.py:158
        color, radii, depth, alpha, _ = self.dgr.GaussianRasterizer(self.settings)(
            means3D=out.xyz, means2D=means2D, opacities=out.opacity, shs=out.features, scales=out.scaling,
            rotations=out.rotation)
        return color, radii, means2D

    def iteration(self, gt, step_optimizer=True):
        color, radii, means2D = self.render()
        if self.name == "ours":
            loss = loss_utils.l1_ssim_loss(color, gt, 0.2)
        else:
            loss = torch_loss(color, gt, 0.2)
        loss.backward()
        if self.name == "ours":
            optim.update_densification_stats(self.stats, radii, means2D.grad)
        else:  # street_gaussian_model.py:555-578
            off, vis_all, rf, g = 0, radii > 0, radii.float(), means2D.grad
            for s, n in zip(self.stats, sizes):
                vis, gg = vis_all[off:off + n], g[off:off + n]
                s.max_radii2D[vis] = torch.max(s.max_radii2D[vis], rf[off:off + n][vis])
                s.xyz_gradient_accum[vis, 0:1] += torch.norm(gg[vis, :2], dim=-1, keepdim=True)
                s.xyz_gradient_accum[vis, 1:2] += torch.norm(gg[vis, 2:], dim=-1, keepdim=True)
                s.denom[vis] += 1
                off += n
        if step_optimizer:
            if self.name == "ours":
                optim.fused_adam_step(self.opts)
            else:
                for o in self.opts:
                    o.step()
            for o in self.opts:
                o.zero_grad(set_to_none=True)
            self.rots.grad = self.trans.grad = None
        return float(loss.item())  # train.py reads the loss every iteration for its running average


def timed(arm, gt, iters, warmup):
    for _ in range(warmup):
        arm.iteration(gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = arm.iteration(gt)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, loss


import diff_gaussian_rasterization as ours_dgr
ref_dgr = load_ref()
ours = Arm("ours", ours_dgr)
# ground truth: the render of a perturbed copy of the scene
with torch.no_grad():
    pert = Arm("ours", ours_dgr)
    for p in pert.subs:
        p["features_dc"].add_(torch.randn_like(p["features_dc"]) * 0.1)
        p["xyz"].add_(torch.randn_like(p["xyz"]) * 0.01)
    gt = pert.render()[0].clamp(0, 1).detach()
    del pert
out = {"config": {"workload": f"training iteration, street scene {P} Gaussians as 1 background + {len(actors)} actors, "
                              f"{sc.width}x{sc.height}, L1 + 0.2 DSSIM, statistics + Adam", "iters": args.iters, "warmup": args.warmup}}
if ref_dgr is not None:
    ref = Arm("reference", ref_dgr)
    # gradient match of the first iteration (no optimiser step)
    ours.iteration(gt, step_optimizer=False); ref.iteration(gt, step_optimizer=False)
    worst = {}
    for po, pr in zip(ours.subs, ref.subs):
        for k in ORDER:
            a, b = po[k].grad, pr[k].grad
            e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            worst[k] = max(worst.get(k, 0.0), e)
    worst["obj_rots"] = float((ours.rots.grad - ref.rots.grad).abs().max() / ref.rots.grad.abs().max())
    worst["obj_trans"] = float((ours.trans.grad - ref.trans.grad).abs().max() / ref.trans.grad.abs().max())
    out["grad_match_iter0_rel"] = {k: round(v, 7) for k, v in worst.items()}
    for arm in (ours, ref):
        for o in arm.opts:
            o.zero_grad(set_to_none=True)
        arm.rots.grad = arm.trans.grad = None
    ms_ref, loss_ref = timed(ref, gt, max(5, args.iters // 3), 2)
    out["reference"] = {"ms_per_iter": round(ms_ref, 3), "iters_per_s": round(1000.0 / ms_ref, 2), "last_loss": loss_ref}
    del ref
    torch.cuda.empty_cache()
ms, loss = timed(ours, gt, args.iters, args.warmup)
out["ours"] = {"ms_per_iter": round(ms, 3), "iters_per_s": round(1000.0 / ms, 2), "last_loss": loss}
if "reference" in out:
    out["speedup"] = round(out["reference"]["ms_per_iter"] / ms, 2)
print(json.dumps(out))
