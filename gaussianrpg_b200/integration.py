"""Drop-in installation of the fused "next row" operators into the reference's own classes.

The rasterizer needs no installation -- the package `diff_gaussian_rasterization` at the repository root is the
drop-in.  The scene-graph compose and the optimiser plumbing live *inside* the reference's
`lib/models/street_gaussian_model.py`, so this module patches that class in place; the callers
(`lib/models/street_gaussian_renderer.py:166-196`, `train.py:277-281,305-307`) stay as they are:

    import lib.models.street_gaussian_model as sgm
    from gaussianrpg_b200 import integration
    integration.install(sgm.StreetGaussianModel)

* `get_xyz / get_rotation / get_scaling / get_opacity / get_features` (street_gaussian_model.py:295-384,438-453) are
  served from ONE fused `compose_scene` call per `parse_camera` (computed lazily on the first getter access);
* `update_optimizer` (:536-553) steps every sub-model's Adam in one launch (`fused_adam_step`), then lets the
  non-Gaussian optimisers (actor pose, sky cube map, colour / pose correction) step as before;
* `set_max_radii2D` and `add_densification_stats` (:555-578) become one launch each over all sub-models (instead of
  four masked read-modify-writes per sub-model), with the caller's visibility mask.

Everything the reference computes in `parse_camera` is reused as is: the per-actor pose is read back out of the
expanded `self.obj_rots / self.obj_trans` rows (one row per actor, autograd intact), the flip mask out of
`self.flip_mask`.  Configurations the fused path does not cover (`use_pose_correction`, a background mask) fall back
to the original getters.  The operators are injectable so that the plumbing can be tested without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional


def _fourier_time(actor, frame) -> float:
    """gaussian_model_actor.py:74-75"""
    return actor.fourier_scale * ((frame - actor.start_frame) / (actor.end_frame - actor.start_frame))


def install(cls, compose: Optional[Callable] = None, idft_base: Optional[Callable] = None,
            adam_step: Optional[Callable] = None, stats_update: Optional[Callable] = None):
    """Patch `cls` (the reference's StreetGaussianModel).  Returns a dict with the original attributes (pass it to
    `uninstall`).  `compose`, `idft_base`, `adam_step`, `stats_update` default to the CUDA operators of this package."""
    from . import scene_compose, optim
    compose = compose or scene_compose.compose_scene
    idft_base = idft_base or scene_compose.idft_base
    adam_step = adam_step or optim.fused_adam_step
    stats_update = stats_update or optim.update_densification_stats
    SubModel, DensifyStats = scene_compose.SubModel, optim.DensifyStats
    names = ("parse_camera", "get_xyz", "get_rotation", "get_scaling", "get_opacity", "get_features",
             "update_optimizer", "set_max_radii2D", "add_densification_stats")
    orig = {n: cls.__dict__[n] for n in names}

    def sub(m):
        return SubModel(m._xyz, m._scaling, m._rotation, m._opacity, m._features_dc, m._features_rest)

    def fused_ok(self):
        if getattr(self, "use_pose_correction", False):
            return False
        if self.get_visibility('background') and getattr(self.background, "background_mask", None) is not None:
            return False
        return True

    def parse_camera(self, camera):
        orig["parse_camera"](self, camera)
        self._grpg_composed = None  # recomputed on the first getter access for this camera

    def composed(self):
        c = getattr(self, "_grpg_composed", None)
        if c is not None:
            return c
        bk = sub(self.background) if self.get_visibility('background') else None
        actors = [getattr(self, n) for n in self.graph_obj_list]
        rots = trans = idft = flips = None
        if actors:
            # parse_camera (:265-281) expanded the K poses to one row per Gaussian; take one row per actor back
            sizes = [a.get_xyz.shape[0] for a in actors]
            starts, off = [], 0
            for n in sizes:
                starts.append(min(off, self.obj_rots.shape[0] - 1))
                off += n
            rots, trans = self.obj_rots[starts], self.obj_trans[starts]
            idft = [idft_base(_fourier_time(a, self.frame), a.fourier_dim) for a in actors]
            flips, off = [], 0
            for n in sizes:
                flips.append(self.flip_mask[off:off + n])
                off += n
        self._grpg_composed = compose(bk, [sub(a) for a in actors], rots, trans, idft, flips)
        return self._grpg_composed

    def getter(field, name):
        def fget(self):
            if not fused_ok(self):
                return orig[name].fget(self)
            return getattr(composed(self), field)
        return property(fget)

    def update_optimizer(self, exclude_list=[]):
        opts = [getattr(self, n).optimizer for n in self.model_name_id.keys()
                if not any(n.startswith(e) for e in exclude_list)]
        adam_step(opts)
        self._grpg_composed = None  # the parameters changed in place: the cached compose is stale
        for o in opts:
            o.zero_grad(set_to_none=True)
        for extra in ("actor_pose", "sky_cubemap", "color_correction", "pose_correction"):
            m = getattr(self, extra, None)
            if m is not None:
                m.update_optimizer()

    def _stats(self):
        return [DensifyStats(m.max_radii2D, m.xyz_gradient_accum, m.denom)
                for m in (getattr(self, n) for n in self.graph_gaussian_range.keys())]

    # The reference's two calls keep their own semantics -- each applies its update immediately, with the mask the
    # caller passes (street_gaussian_model.py:555-578) -- as one launch each over all sub-models.
    def set_max_radii2D(self, radii, visibility_filter):
        stats_update(_stats(self), radii, None, visibility_filter, 1)

    def add_densification_stats(self, viewspace_point_tensor, visibility_filter):
        stats_update(_stats(self), None, viewspace_point_tensor.grad, visibility_filter, 2)

    # anything that replaces or rewrites parameter tensors (densification, opacity reset) or changes which sub-models
    # are composed invalidates the per-camera cache
    def invalidating(name):
        fn = cls.__dict__[name]

        def wrapped(self, *a, **kw):
            self._grpg_composed = None
            try:
                return fn(self, *a, **kw)
            finally:
                self._grpg_composed = None
        wrapped.__name__ = name
        return wrapped

    for name in ("densify_and_prune", "reset_opacity", "set_visibility"):
        if name in cls.__dict__ and callable(cls.__dict__[name]):
            orig[name] = cls.__dict__[name]
            setattr(cls, name, invalidating(name))

    cls.parse_camera = parse_camera
    cls.get_xyz = getter("xyz", "get_xyz")
    cls.get_rotation = getter("rotation", "get_rotation")
    cls.get_scaling = getter("scaling", "get_scaling")
    cls.get_opacity = getter("opacity", "get_opacity")
    cls.get_features = getter("features", "get_features")
    cls.update_optimizer = update_optimizer
    cls.set_max_radii2D = set_max_radii2D
    cls.add_densification_stats = add_densification_stats
    return orig


def uninstall(cls, orig) -> None:
    for n, v in orig.items():
        setattr(cls, n, v)
