"""Seeded synthetic inputs for the parity tests and bench.py (no dataset or network needed).

Camera conventions follow the reference callers: `viewmatrix` is world_view_transform and
`projmatrix` is full_proj_transform = (W2C^T-stored) view @ projection^T, i.e. both are passed
transposed (lib/utils/camera_utils.py:50-58); projection matrices restate
lib/utils/graphics_utils.py:51-70 (getProjectionMatrix) and :72-94 (getProjectionMatrixK).
All tensors are generated on the CPU in float32 from a torch.Generator seed so every
implementation sees identical bits.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch


def projection_from_fov(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def projection_from_intrinsics(fx, fy, cx, cy, W, H, znear: float, zfar: float) -> torch.Tensor:
    P = torch.zeros(4, 4)
    P[0, 0] = 2 * fx / W
    P[0, 2] = -1 + 2 * (cx / W)
    P[1, 1] = 2 * fy / H
    P[1, 2] = -1 + 2 * (cy / H)
    P[2, 2] = (zfar + znear) / (zfar - znear)
    P[2, 3] = -2 * zfar * znear / (zfar - znear)
    P[3, 2] = 1.0
    return P


@dataclass
class Scene:
    """Everything one rasterizer call needs; tensors live on the CPU until .to(device)."""
    width: int
    height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    campos: torch.Tensor
    bg: torch.Tensor
    means3D: torch.Tensor
    opacities: torch.Tensor
    scales: Optional[torch.Tensor] = None
    rotations: Optional[torch.Tensor] = None
    shs: Optional[torch.Tensor] = None
    sh_degree: int = 0
    colors_precomp: Optional[torch.Tensor] = None
    cov3D_precomp: Optional[torch.Tensor] = None
    semantics: Optional[torch.Tensor] = None
    scale_modifier: float = 1.0
    name: str = "scene"
    extra: Dict = field(default_factory=dict)

    def to(self, device) -> "Scene":
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Scene(**kw)

    def settings(self, debug: bool = False):
        from .rasterizer import GaussianRasterizationSettings
        return GaussianRasterizationSettings(
            image_height=self.height, image_width=self.width, tanfovx=self.tanfovx, tanfovy=self.tanfovy, bg=self.bg,
            scale_modifier=self.scale_modifier, viewmatrix=self.viewmatrix, projmatrix=self.projmatrix,
            sh_degree=self.sh_degree, campos=self.campos, prefiltered=False, debug=debug)

    def raster_kwargs(self):
        return dict(means3D=self.means3D, opacities=self.opacities, shs=self.shs, colors_precomp=self.colors_precomp,
                    scales=self.scales, rotations=self.rotations, cov3D_precomp=self.cov3D_precomp,
                    semantics=self.semantics)


def _camera(view_w2c: torch.Tensor, proj: torch.Tensor):
    """(viewmatrix, projmatrix, campos) as camera_utils.py:50-58 builds them."""
    world_view = view_w2c.t().contiguous()
    full_proj = (world_view.unsqueeze(0).bmm(proj.t().unsqueeze(0))).squeeze(0).contiguous()
    campos = world_view.inverse()[3, :3].contiguous()
    return world_view.float(), full_proj.float(), campos.float()


def plumbing_scene(P: int = 128, W: int = 64, H: int = 64, S: int = 0, white_bg: bool = False, sh_degree: int = 1,
                   seed: int = 0) -> Scene:
    """BASELINE config #1: camera at the origin looking +z, fov 60 deg, some points behind the
    near plane, anisotropic scales, un-normalised quaternions."""
    g = torch.Generator().manual_seed(seed)
    fov = math.radians(60.0)
    proj = projection_from_fov(0.001, 1000.0, fov, fov)
    view, full, campos = _camera(torch.eye(4), proj)
    means = torch.empty(P, 3)
    means[:, :2] = torch.rand(P, 2, generator=g) * 3.0 - 1.5
    means[:, 2] = torch.rand(P, generator=g) * 5.9 + 0.1
    scales = torch.exp(torch.rand(P, 3, generator=g) * (math.log(0.5) - math.log(0.03)) + math.log(0.03))
    rots = torch.randn(P, 4, generator=g)
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) * 1.5)
    M = 16 if sh_degree == 3 else (9 if sh_degree == 2 else 4)
    shs = torch.randn(P, M, 3, generator=g) * 0.2
    shs[:, 0] = torch.rand(P, 3, generator=g) * 2 - 1
    sem = torch.rand(P, S, generator=g) if S > 0 else None
    return Scene(width=W, height=H, tanfovx=math.tan(fov / 2), tanfovy=math.tan(fov / 2), viewmatrix=view,
                 projmatrix=full, campos=campos, bg=torch.ones(3) if white_bg else torch.zeros(3), means3D=means,
                 opacities=opac, scales=scales, rotations=rots, shs=shs, sh_degree=sh_degree, semantics=sem,
                 name=f"plumbing_P{P}_{W}x{H}_S{S}")


def test_script_scene(P: int = 10000, W: int = 1242, H: int = 375, S: int = 0, seed: int = 0) -> Scene:
    """BASELINE config #2: the distributions and camera literals of script/test_gaussian_rasterization.py
    (:7-19 camera, :44-52 inputs; sh_degree 0 with [N,4,3] SH, rotations[:,0]=1)."""
    g = torch.Generator().manual_seed(seed)
    view = torch.tensor([[0.9598, 0.0081, 0.2806, 0.0], [-0.0123, 0.9998, 0.0134, 0.0], [-0.2804, -0.0163, 0.9597, 0.0],
                         [-2.0954, -0.0935, 4.9320, 1.0]])
    full = torch.tensor([[1.1205, 0.0312, 0.2806, 0.2806], [-0.0144, 3.8661, 0.0134, 0.0134],
                         [-0.3274, -0.0632, 0.9598, 0.9597], [-2.4464, -0.3614, 4.9225, 4.9320]])
    campos = torch.tensor([6.2808e-01, 1.4572e-03, -5.3226e+00])
    rots = torch.rand(P, 4, generator=g)
    rots[:, 0] = 1
    return Scene(width=W, height=H, tanfovx=math.tan(1.416 * 0.5), tanfovy=math.tan(0.506 * 0.5), viewmatrix=view,
                 projmatrix=full, campos=campos, bg=torch.zeros(3), means3D=torch.rand(P, 3, generator=g),
                 opacities=torch.rand(P, 1, generator=g), scales=torch.rand(P, 3, generator=g), rotations=rots,
                 shs=torch.rand(P, 4, 3, generator=g), sh_degree=0,
                 semantics=torch.rand(P, S, generator=g) if S > 0 else None, name=f"testscript_P{P}_{W}x{H}_S{S}")


def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ar, ax, ay, az = a.unbind(-1)
    br, bx, by, bz = b.unbind(-1)
    return torch.stack([ar * br - ax * bx - ay * by - az * bz, ar * bx + ax * br + ay * bz - az * by,
                        ar * by - ax * bz + ay * br + az * bx, ar * bz + ax * by - ay * bx + az * br], -1)


def street_scene(P: int = 2_000_000, W: int = 1920, H: int = 1280, n_actors: int = 8, actor_points: int = 20_000,
                 sh_degree: int = 1, seed: int = 0, cam_z: float = 0.0) -> Scene:
    """BASELINE config #3: Waymo-shaped street scene: background = 40% ground plane, 40% facades,
    20% box-uniform; `n_actors` vehicles of `actor_points` Gaussians composed background-first the way
    lib/models/street_gaussian_model.py:314-367 does (quaternion product + normalise, rotate + translate).
    Camera: fx=fy=2083.09 px, principal point at the centre, y-down / z-forward at (0,0,cam_z)."""
    g = torch.Generator().manual_seed(seed)
    n_act = n_actors * actor_points if P >= 4 * n_actors * actor_points else 0
    n_bg = P - n_act
    n_ground, n_facade = int(0.4 * n_bg), int(0.4 * n_bg)
    n_box = n_bg - n_ground - n_facade
    U = lambda n, lo, hi: torch.rand(n, generator=g) * (hi - lo) + lo  # noqa: E731
    ground = torch.stack([U(n_ground, -20, 20), 1.8 + torch.randn(n_ground, generator=g) * 0.05, U(n_ground, 1, 120)], 1)
    side = torch.where(torch.rand(n_facade, generator=g) < 0.5, -1.0, 1.0)
    facade = torch.stack([side * (12 + U(n_facade, 0, 3)), U(n_facade, -15, 1.8), U(n_facade, 1, 150)], 1)
    box = torch.stack([U(n_box, -25, 25), U(n_box, -15, 2), U(n_box, 1, 150)], 1)
    means = [ground, facade, box]
    scales_bg = torch.exp(torch.randn(n_bg, 3, generator=g) * 0.6 + math.log(0.08))
    thin = torch.randint(0, 3, (n_bg,), generator=g)
    scales_bg[torch.arange(n_bg), thin] *= 0.2
    rots_bg = torch.nn.functional.normalize(torch.randn(n_bg, 4, generator=g), dim=1)
    scales, rots = [scales_bg], [rots_bg]
    graph = []  # per actor: (local xyz, local quaternion, pose quaternion, pose translation) -- see street_scene_graph
    if n_act:
        lanes = [-1.75, 1.75, -5.25, 5.25]
        for a in range(n_actors):
            local = torch.stack([U(actor_points, -2.25, 2.25), U(actor_points, -0.8, 0.8), U(actor_points, -0.95, 0.95)], 1)
            yaw = float(U(1, -math.pi, math.pi))
            q_obj = torch.tensor([math.cos(yaw / 2), 0.0, math.sin(yaw / 2), 0.0])  # rotation about the y (up) axis
            Rm = torch.tensor([[math.cos(yaw), 0.0, math.sin(yaw)], [0.0, 1.0, 0.0], [-math.sin(yaw), 0.0, math.cos(yaw)]])
            trans = torch.tensor([lanes[a % 4], 1.0, 8.0 + 6.0 * a])
            means.append(local @ Rm.t() + trans)
            q_local = torch.nn.functional.normalize(torch.randn(actor_points, 4, generator=g), dim=1)
            rots.append(torch.nn.functional.normalize(_quat_mul(q_obj.expand_as(q_local), q_local), dim=1))
            graph.append((local, q_local, q_obj, trans))
            scales.append(torch.exp(torch.randn(actor_points, 3, generator=g) * 0.5 + math.log(0.05)))
    means3D = torch.cat(means, 0).float().contiguous()
    P_tot = means3D.shape[0]
    opac = torch.sigmoid(torch.randn(P_tot, 1, generator=g) * 2.0 + 1.0)
    M = (sh_degree + 1) ** 2
    shs = torch.randn(P_tot, M, 3, generator=g) * 0.1
    shs[:, 0] = torch.randn(P_tot, 3, generator=g) * 0.5
    f = 2083.09
    proj = projection_from_intrinsics(f, f, W / 2, H / 2, W, H, 0.001, 1000.0)
    w2c = torch.eye(4)
    w2c[2, 3] = -cam_z  # camera at z = cam_z looking down +z
    view, full, campos = _camera(w2c, proj)
    return Scene(width=W, height=H, tanfovx=W / (2 * f), tanfovy=H / (2 * f), viewmatrix=view, projmatrix=full,
                 campos=campos, bg=torch.zeros(3), means3D=means3D, opacities=opac,
                 scales=torch.cat(scales, 0).float().contiguous(), rotations=torch.cat(rots, 0).float().contiguous(),
                 shs=shs.contiguous(), sh_degree=sh_degree, name=f"street_P{P_tot}_{W}x{H}",
                 extra={"n_bkgd": n_bg, "graph": graph})


def street_scene_graph(sc: Scene, fourier_dim: int = 5, time: float = 0.3, seed: int = 1):
    """The sub-model parameters behind `street_scene` in the reference's storage format (gaussian_model.py:207-251:
    raw xyz, log-scales, raw quaternions, opacity logits, SH dc / rest; actors with `fourier_dim` dc rows,
    gaussian_model_actor.py:73-82) plus the per-actor pose, such that composing them
    (street_gaussian_model.py:295-384) reproduces the scene's rasterizer inputs.  Returns
    (background dict, [actor dicts], obj_rots [K,4], obj_trans [K,3], idft [K][F])."""
    from .scene_compose import idft_base
    g = torch.Generator().manual_seed(seed)
    n_bg, graph = sc.extra["n_bkgd"], sc.extra["graph"]

    def sub(sl, xyz, rot, F):
        n = xyz.shape[0]
        dc = torch.zeros(n, F, 3)
        base = torch.tensor(idft_base(time, F))
        dc[:, 0] = sc.shs[sl, 0]  # idft[0] = cos(0) = 1: the composed dc equals the scene's when the other rows are 0
        if F > 1:
            extra = torch.randn(n, F - 1, 3, generator=g) * 0.05
            dc[:, 1:] = extra
            dc[:, 0] -= (extra * base[1:, None]).sum(1)
        o = sc.opacities[sl].clamp(1e-6, 1 - 1e-6)
        return dict(xyz=xyz.contiguous(), scaling=torch.log(sc.scales[sl]).contiguous(), rotation=rot.contiguous(),
                    opacity=torch.log(o / (1 - o)).contiguous(), features_dc=dc.contiguous(),
                    features_rest=sc.shs[sl, 1:].contiguous())

    bk = sub(slice(0, n_bg), sc.means3D[:n_bg], sc.rotations[:n_bg], 1)
    actors, off = [], n_bg
    for local, q_local, _, _ in graph:
        n = local.shape[0]
        actors.append(sub(slice(off, off + n), local, q_local, fourier_dim))
        off += n
    K = len(graph)
    obj_rots = torch.stack([q for _, _, q, _ in graph]) if K else torch.zeros(0, 4)
    obj_trans = torch.stack([t for _, _, _, t in graph]) if K else torch.zeros(0, 3)
    return bk, actors, obj_rots, obj_trans, [idft_base(time, fourier_dim) for _ in range(K)]
