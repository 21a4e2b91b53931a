"""BASELINE config #4 harness: the reference's TRAINING LOOP around the op (train.py:64-307), with densification.

train.py itself cannot run in this image (plyfile, bidict, nvdiffrast, simple_knn are absent), so the loop is restated
here around the operators, on the street scene stored as sub-models (1 background + K actors), as two arms:

  ours      : compose_scene + GaussianRasterizer + l1_ssim_loss + update_densification_stats + fused_adam_step
  reference : the reference's formulation of every stage -- stock PyTorch ops for compose / loss / statistics, one
              torch.optim.Adam(eps=1e-15) per sub-model, and the UNMODIFIED reference rasterizer (oracle/_ref).

Per iteration (train.py:110-118,229,277-307): pick a camera, compose (street_gaussian_model.py:295-384), render,
L1 + lambda_dssim * DSSIM, backward, `set_max_radii2D` + `add_densification_stats` (:555-578), every
`densification_interval` iterations `densify_and_prune` per sub-model, then the optimiser step.

Densify / clone / split / prune and the optimiser-state surgery restate lib/models/gaussian_model.py:366-553 with the
per-model rules of gaussian_model_bkgd.py:73-112 (abs-gradient statistic, `densify_grad_threshold_bkgd`) and
gaussian_model_actor.py:206-263 (`densify_grad_threshold_obj`) and the thresholds of
configs/example/waymo_train_002.yaml:34-57, in plain torch, shared by BOTH arms: the arms differ only in the operators
under test.  The split's `torch.normal` draws come from a generator seeded per (iteration, sub-model), so both arms see
the same samples as long as they select the same Gaussians.
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from gaussianrpg_b200 import loss_utils, optim, scene_compose, synthetic  # noqa: E402

ORDER = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
GROUP = dict(xyz="xyz", features_dc="f_dc", features_rest="f_rest", opacity="opacity", scaling="scaling",
             rotation="rotation", semantic="semantic")  # parameter -> optimiser group name (gaussian_model.py:292-304)
PARAMS = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation", "semantic")

# configs/example/waymo_train_002.yaml:34-57
CFG = dict(densification_interval=100, densify_from_iter=500, densify_until_iter=25000,
           densify_grad_threshold_bkgd=0.0006, densify_grad_abs_bkgd=True, densify_grad_threshold_obj=0.0002,
           densify_grad_abs_obj=False, densify_grad_threshold=0.0002, min_opacity=0.005, opacity_reset_interval=3000,
           percent_dense=0.01, percent_big_ws=0.1, position_lr_init=0.00016, feature_lr=0.0025, opacity_lr=0.05,
           scaling_lr=0.005, rotation_lr=0.001, semantic_lr=0.01, lambda_dssim=0.2)


def load_reference_extension():
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_ref", ROOT / "oracle" / "build_ref.py")
    build_ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build_ref)
    return build_ref.load() if build_ref.available() else None


def quaternion_to_matrix(r):
    """lib/utils/general_utils.py:125-146"""
    q = r / torch.sqrt((r * r).sum(1))[:, None]
    w, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)


def torch_reference_ops():
    """The reference's formulation of compose and loss in stock PyTorch (shared with the operator tests)."""
    from test_compose_gpu import _torch_reference as torch_compose
    from test_loss_gpu import _torch_reference as torch_loss
    return torch_compose, torch_loss


class Cameras:
    """A few poses along the street (poses along z) with their ground-truth images."""

    def __init__(self, points, width, height, n_cams, dev, n_actors=8, actor_points=20_000):
        self.scenes = [synthetic.street_scene(P=points, W=width, H=height, n_actors=n_actors, actor_points=actor_points,
                                              cam_z=float(i)) for i in range(n_cams)]
        self.dev = dev
        self.gt = [None] * n_cams

    def settings(self, dgr, i):
        sc = self.scenes[i]
        d = self.dev
        return dgr.GaussianRasterizationSettings(
            image_height=sc.height, image_width=sc.width, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=sc.bg.to(d),
            scale_modifier=1.0, viewmatrix=sc.viewmatrix.to(d), projmatrix=sc.projmatrix.to(d), sh_degree=sc.sh_degree,
            campos=sc.campos.to(d), prefiltered=False, debug=False)


class Arm:
    def __init__(self, name, dgr, cams: Cameras, dev, cfg=None, bkgd_extent=20.0, actor_extent=5.0):
        self.name, self.dgr, self.cams, self.dev = name, dgr, cams, dev
        self.cfg = dict(CFG, **(cfg or {}))
        sc = cams.scenes[0]
        bk, actors, rots0, trans0, self.idft = synthetic.street_scene_graph(sc)
        self.subs = []
        for d in [bk] + actors:
            p = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in d.items()}
            p["semantic"] = torch.nn.Parameter(torch.zeros(d["xyz"].shape[0], 0, device=dev))
            self.subs.append(p)
        self.is_bkgd = [True] + [False] * len(actors)
        self.extent = [bkgd_extent] + [actor_extent] * len(actors)  # scene_radius / actor extent
        self.rots = rots0.to(dev).requires_grad_(True)   # actor_pose parameters stay on their own optimiser in the
        self.trans = trans0.to(dev).requires_grad_(True)  # reference (actor_pose.py); not stepped here
        self.opts = [self._training_setup(p) for p in self.subs]
        self.stats = [[torch.zeros(n, device=dev), torch.zeros(n, 2, device=dev), torch.zeros(n, 1, device=dev)]
                      for n in self.sizes()]
        self._settings = [cams.settings(dgr, i) for i in range(len(cams.scenes))]
        self.torch_compose, self.torch_loss = torch_reference_ops()

    # ---- gaussian_model.py:286-318 --------------------------------------------------------------------------------
    def _training_setup(self, p, spatial_lr_scale=3.0):
        c = self.cfg
        groups = [{'params': [p["xyz"]], 'lr': c["position_lr_init"] * spatial_lr_scale, "name": "xyz"},
                  {'params': [p["features_dc"]], 'lr': c["feature_lr"], "name": "f_dc"},
                  {'params': [p["features_rest"]], 'lr': c["feature_lr"] / 20.0, "name": "f_rest"},
                  {'params': [p["opacity"]], 'lr': c["opacity_lr"], "name": "opacity"},
                  {'params': [p["scaling"]], 'lr': c["scaling_lr"], "name": "scaling"},
                  {'params': [p["rotation"]], 'lr': c["rotation_lr"], "name": "rotation"},
                  {'params': [p["semantic"]], 'lr': c["semantic_lr"], "name": "semantic"}]
        cls = optim.FusedAdam if self.name == "ours" else torch.optim.Adam
        return cls(groups, lr=0.0, eps=1e-15)

    def sizes(self):
        return [int(p["xyz"].shape[0]) for p in self.subs]

    def total(self):
        return sum(self.sizes())

    # ---- one iteration ----------------------------------------------------------------------------------------------
    def compose(self):
        subm = [scene_compose.SubModel(*(p[k] for k in ORDER)) for p in self.subs]
        if self.name == "ours":
            return scene_compose.compose_scene(subm[0], subm[1:], self.rots, self.trans, self.idft, None)
        flips = [torch.zeros(n, dtype=torch.bool, device=self.dev) for n in self.sizes()[1:]]
        return self.torch_compose(subm[0], subm[1:], self.rots, self.trans, self.idft, flips)

    def render(self, cam):
        out = self.compose()
        P = out.xyz.shape[0]
        means2D = torch.zeros(P, 3, device=self.dev, requires_grad=True)  # street_gaussian_renderer.py:158
        color, radii, depth, alpha, _ = self.dgr.GaussianRasterizer(self._settings[cam])(
            means3D=out.xyz, means2D=means2D, opacities=out.opacity, shs=out.features, scales=out.scaling,
            rotations=out.rotation)
        return color, radii, means2D

    def iteration(self, cam, gt, step_optimizer=True, densify_seed=None, forced=None):
        """One training iteration; `densify_seed` not None = this iteration densifies and prunes (train.py:283-291).
        `forced` = the other arm's `last_masks` (lock-step runs, see densify_and_prune)."""
        color, radii, means2D = self.render(cam)
        lam = self.cfg["lambda_dssim"]
        loss = loss_utils.l1_ssim_loss(color, gt, lam) if self.name == "ours" else self.torch_loss(color, gt, lam)
        loss.backward()
        sizes = self.sizes()
        if self.name == "ours":
            optim.update_densification_stats([optim.DensifyStats(*s) for s in self.stats], radii, means2D.grad)
        else:  # street_gaussian_model.py:555-578
            off, vis_all, rf, g = 0, radii > 0, radii.float(), means2D.grad
            for s, n in zip(self.stats, sizes):
                vis, gg = vis_all[off:off + n], g[off:off + n]
                s[0][vis] = torch.max(s[0][vis], rf[off:off + n][vis])
                s[1][vis, 0:1] += torch.norm(gg[vis, :2], dim=-1, keepdim=True)
                s[1][vis, 1:2] += torch.norm(gg[vis, 2:], dim=-1, keepdim=True)
                s[2][vis] += 1
                off += n
        if densify_seed is not None:
            self.densify_and_prune(densify_seed, forced)
        if step_optimizer:
            if self.name == "ours":
                optim.fused_adam_step(self.opts)
            else:
                for o in self.opts:
                    o.step()
            for o in self.opts:
                o.zero_grad(set_to_none=True)
            self.rots.grad = self.trans.grad = None
        return loss

    # ---- optimiser-state surgery, gaussian_model.py:366-411 --------------------------------------------------------
    def _prune_points(self, k, keep):
        opt, p = self.opts[k], self.subs[k]
        for group in opt.param_groups:
            old = group["params"][0]
            stored = opt.state.get(old, None)
            new = torch.nn.Parameter(old[keep].requires_grad_(True))
            if stored is not None:
                stored["exp_avg"] = stored["exp_avg"][keep]
                stored["exp_avg_sq"] = stored["exp_avg_sq"][keep]
                del opt.state[old]
                opt.state[new] = stored
            group["params"][0] = new
            p[[n for n, g in GROUP.items() if g == group["name"]][0]] = new
        self.stats[k] = [s[keep] for s in self.stats[k]]

    def _cat_points(self, k, ext):
        opt, p = self.opts[k], self.subs[k]
        for group in opt.param_groups:
            name = [n for n, g in GROUP.items() if g == group["name"]][0]
            old, e = group["params"][0], ext[name]
            stored = opt.state.get(old, None)
            new = torch.nn.Parameter(torch.cat((old, e), 0).requires_grad_(True))
            if stored is not None:
                stored["exp_avg"] = torch.cat((stored["exp_avg"], torch.zeros_like(e)), 0)
                stored["exp_avg_sq"] = torch.cat((stored["exp_avg_sq"], torch.zeros_like(e)), 0)
                del opt.state[old]
                opt.state[new] = stored
            group["params"][0] = new
            p[name] = new
        n_new = ext["xyz"].shape[0]
        z = lambda *s: torch.zeros(*s, device=self.dev)  # noqa: E731
        self.stats[k] = [torch.cat([self.stats[k][0], z(n_new)]), torch.cat([self.stats[k][1], z(n_new, 2)]),
                         torch.cat([self.stats[k][2], z(n_new, 1)])]

    # ---- gaussian_model.py:453-553 with the bkgd / actor rules -----------------------------------------------------
    @torch.no_grad()
    def densify_and_prune(self, seed, forced=None):
        """Every sub-model's clone / split / prune decisions are computed from THIS arm's statistics and parameters
        and kept in `last_masks`.  With `forced` (the `last_masks` of the other arm of a lock-step run, which holds the
        same point set) the other arm's decisions are the ones applied, and the report counts the Gaussians on which
        the two arms decided differently (`disagree`): the arms then keep identical point sets through every event, so
        their trajectories stay comparable iteration by iteration instead of drifting apart after the first
        borderline Gaussian."""
        c = self.cfg
        report, masks = [], []
        for k in range(len(self.subs)):
            p = self.subs[k]
            if self.is_bkgd[k]:
                max_grad, use_abs = c["densify_grad_threshold_bkgd"], c["densify_grad_abs_bkgd"]
            else:
                max_grad, use_abs = c["densify_grad_threshold_obj"], c["densify_grad_abs_obj"]
            accum, denom = self.stats[k][1], self.stats[k][2]
            grads = (accum[:, 1:2] if use_abs else accum[:, 0:1]) / denom
            grads[grads.isnan()] = 0.0
            extent = self.extent[k]
            n0 = p["xyz"].shape[0]
            # clone (:501-528)
            sel = (torch.norm(grads, dim=-1) >= max_grad) & (torch.exp(p["scaling"]).max(1).values <= c["percent_dense"] * extent)
            own = dict(clone=sel)
            if forced is not None:
                sel = forced[k]["clone"]
            n_clone = int(sel.sum())
            self._cat_points(k, {n: p[n][sel] for n in PARAMS})
            p = self.subs[k]
            # split (:453-499)
            n_init, N = p["xyz"].shape[0], 2
            padded = torch.zeros(n_init, device=self.dev)
            padded[:grads.shape[0]] = grads.squeeze(1)
            sel = (padded >= max_grad) & (torch.exp(p["scaling"]).max(1).values > c["percent_dense"] * extent)
            own["split"] = sel
            if forced is not None:
                sel = forced[k]["split"]
            n_split = int(sel.sum())
            gen = torch.Generator(device=self.dev)
            gen.manual_seed(int(seed) * 1000 + k)
            stds = torch.exp(p["scaling"][sel]).repeat(N, 1)
            samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=gen) if n_split else stds
            rots = quaternion_to_matrix(p["rotation"][sel]).repeat(N, 1, 1)
            ext = dict(xyz=torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + p["xyz"][sel].repeat(N, 1),
                       scaling=torch.log(torch.exp(p["scaling"][sel]).repeat(N, 1) / (0.8 * N)),
                       rotation=p["rotation"][sel].repeat(N, 1), features_dc=p["features_dc"][sel].repeat(N, 1, 1),
                       features_rest=p["features_rest"][sel].repeat(N, 1, 1), opacity=p["opacity"][sel].repeat(N, 1),
                       semantic=p["semantic"][sel].repeat(N, 1))
            self._cat_points(k, ext)
            self._prune_points(k, ~torch.cat((sel, torch.zeros(N * n_split, device=self.dev, dtype=torch.bool))))
            p = self.subs[k]
            # prune (:540-547; prune_big_points is False before opacity_reset_interval, train.py:281)
            prune = (torch.sigmoid(p["opacity"]) < c["min_opacity"]).squeeze(1)
            own["prune"] = prune
            if forced is not None:
                prune = forced[k]["prune"]
            n_prune = int(prune.sum())
            self._prune_points(k, ~prune)
            n1 = self.subs[k]["xyz"].shape[0]
            self.stats[k] = [torch.zeros(n1, device=self.dev), torch.zeros(n1, 2, device=self.dev),
                             torch.zeros(n1, 1, device=self.dev)]
            disagree = 0 if forced is None else sum(int((own[m] != forced[k][m]).sum()) for m in ("clone", "split", "prune"))
            report.append(dict(before=n0, cloned=n_clone, split=n_split, pruned=n_prune, after=n1, disagree=disagree))
            masks.append(own)
        self.last_densify, self.last_masks = report, masks
        return report


def make_ground_truth(cams: Cameras, dgr, dev, seed=77):
    """GT = the render of a perturbed copy of the scene, per camera, by this repo's op."""
    with torch.no_grad():
        pert = Arm("ours", dgr, cams, dev)
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        for p in pert.subs:
            p["features_dc"].add_(torch.randn(p["features_dc"].shape, device=dev, generator=g) * 0.1)
            p["xyz"].add_(torch.randn(p["xyz"].shape, device=dev, generator=g) * 0.01)
        for i in range(len(cams.scenes)):
            cams.gt[i] = pert.render(i)[0].clamp(0, 1).detach()
        del pert
    return cams.gt


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)


def run_training(arm: Arm, iters, densify_at, cam_order=None, eval_cam=0, eval_at=(), time_from=None):
    """`iters` iterations cycling over the cameras; densify at the iteration numbers in `densify_at` (1-based, as
    train.py counts).  Returns dict(losses, P_after_densify, renders {it: image}, ms_per_iter over [time_from, iters])."""
    n_cams = len(arm.cams.scenes)
    losses, sizes, renders = [], {}, {}
    e0 = e1 = None
    for it in range(1, iters + 1):
        if time_from is not None and it == time_from:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        cam = (cam_order[it - 1] if cam_order is not None else (it - 1)) % n_cams
        loss = arm.iteration(cam, arm.cams.gt[cam], densify_seed=it if it in densify_at else None)
        losses.append(float(loss.item()))  # train.py reads the loss every iteration for its running average
        if it in densify_at:
            sizes[it] = (arm.sizes(), arm.last_densify)
        if it in eval_at:
            with torch.no_grad():
                renders[it] = arm.render(eval_cam)[0].detach().clone()
    out = dict(losses=losses, sizes=sizes, renders=renders)
    if e0 is not None:
        e1.record()
        torch.cuda.synchronize()
        out["ms_per_iter"] = e0.elapsed_time(e1) / (iters - time_from + 1)
    return out


def run_training_lockstep(leader: Arm, follower: Arm, iters, densify_at, eval_cam=0, eval_at=()):
    """Both arms step through the same iterations side by side.  At a densify event each arm computes its own
    clone / split / prune decisions; the leader's are applied in both (Arm.densify_and_prune, `forced`), and the number
    of Gaussians the follower would have decided differently is recorded.  Returns {arm.name: dict(losses, sizes,
    renders)}; `sizes[it]` of the follower carries `disagree` per sub-model."""
    n_cams = len(leader.cams.scenes)
    out = {a.name: dict(losses=[], sizes={}, renders={}) for a in (leader, follower)}
    for it in range(1, iters + 1):
        cam = (it - 1) % n_cams
        seed = it if it in densify_at else None
        for arm in (leader, follower):
            forced = leader.last_masks if (seed is not None and arm is follower) else None
            loss = arm.iteration(cam, arm.cams.gt[cam], densify_seed=seed, forced=forced)
            o = out[arm.name]
            o["losses"].append(float(loss.item()))
            if seed is not None:
                o["sizes"][it] = (arm.sizes(), arm.last_densify)
            if it in eval_at:
                with torch.no_grad():
                    o["renders"][it] = arm.render(eval_cam)[0].detach().clone()
    return out
