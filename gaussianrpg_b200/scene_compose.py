"""Fused scene-graph compose: sub-model parameters -> the rasterizer's per-Gaussian inputs.

Host-side mirror of what `StreetGaussianModel` computes on every render between `parse_camera` and the rasterizer
call (/root/reference/lib/models/street_gaussian_model.py): `get_xyz` :341-367, `get_rotation` :314-338,
`get_scaling` :296-312, `get_opacity` :438-453, `get_features` :370-384 -- here one CUDA launch forward and one
backward (`grpg_compose_forward/backward`, include/grpg_compose.h) instead of ~40 PyTorch kernels with their
clones, expanded pose rows and concatenations.  Sub-models keep the reference's parameter tensors (`_xyz`,
`_scaling`, `_rotation`, `_opacity`, `_features_dc`, `_features_rest`), so optimisers, densification and checkpoints
are untouched; gradients flow to every parameter and to the per-actor pose (`obj_rots`, `obj_trans`), exactly the
leaves autograd reaches through the reference's getters.  There is no CPU path.

    out = compose_scene(background, actors, obj_rots, obj_trans, idft, flip_masks)
    out.xyz, out.rotation, out.scaling, out.opacity, out.features      # == model.get_xyz, ... of the reference
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence

import numpy as np
import torch

from . import _lib

FLIP_AXIS = 1  # street_gaussian_model.py:58
_N_PARAMS = 6


class SubModel(NamedTuple):
    """Parameters of one `GaussianModel` as the reference stores them (gaussian_model.py:207-251)."""
    xyz: torch.Tensor            # [n,3]
    scaling: torch.Tensor        # [n,3]  log-scales (`_scaling`)
    rotation: torch.Tensor       # [n,4]  raw quaternions (`_rotation`)
    opacity: torch.Tensor        # [n,1]  logits (`_opacity`)
    features_dc: torch.Tensor    # [n,F,3]  F = 1 for the background, fourier_dim for actors
    features_rest: torch.Tensor  # [n,M-1,3]


class ComposedScene(NamedTuple):
    xyz: torch.Tensor        # [P,3]    == StreetGaussianModel.get_xyz
    rotation: torch.Tensor   # [P,4]    == get_rotation
    scaling: torch.Tensor    # [P,3]    == get_scaling
    opacity: torch.Tensor    # [P,1]    == get_opacity
    features: torch.Tensor   # [P,M,3]  == get_features


def idft_base(time: float, dim: int) -> List[float]:
    """`IDFT(time, dim)[0]` of lib/utils/sh_utils.py:120-130: cos(pi t k) for even k, sin(pi t (k+1)) for odd k,
    in float32 like the reference's torch ops."""
    k = np.arange(dim)
    arg = np.float32(np.pi) * np.float32(time) * np.where(k % 2 == 0, k, k + 1).astype(np.float32)
    return np.where(k % 2 == 0, np.cos(arg), np.sin(arg)).astype(np.float32).tolist()


def _contig(t: torch.Tensor) -> torch.Tensor:
    t = t if t.dtype == torch.float32 else t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _descriptors(subs, is_actor, fourier, pose, idft_h, flips, grads=None):
    """`pose` = (obj_rots [K,4], obj_trans [K,3]) contiguous float32 CUDA tensors: the descriptors carry device
    pointers to their rows, so building them needs no device->host read."""
    arr = (_lib.ComposeSubmodel * len(subs))()
    rot_p = pose[0].data_ptr() if pose[0] is not None else 0
    trans_p = pose[1].data_ptr() if pose[1] is not None else 0
    ai = 0
    for k, s in enumerate(subs):
        d = arr[k]
        n = int(s[0].shape[0])
        d.n, d.is_actor, d.fourier_dim = n, int(is_actor[k]), int(fourier[k])
        if n:
            d.xyz, d.scaling, d.rotation, d.opacity, d.features_dc = (t.data_ptr() for t in s[:5])
            d.features_rest = s[5].data_ptr() if s[5].numel() else None
        if is_actor[k]:
            d.obj_rot, d.obj_trans = rot_p + 16 * ai, trans_p + 12 * ai
            for i, v in enumerate(idft_h[ai]):
                d.idft[i] = v
            f = flips[ai]
            d.flip_mask = f.data_ptr() if f is not None and f.numel() else None
            ai += 1
        if grads is not None and n:
            g = grads[k]
            d.d_xyz, d.d_scaling, d.d_rotation, d.d_opacity, d.d_features_dc = (t.data_ptr() for t in g[:5])
            d.d_features_rest = g[5].data_ptr() if g[5].numel() else None
    return arr


class _ComposeScene(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, obj_rots, obj_trans, *params):
        lib = _lib.load()
        n_sub, is_actor, fourier, M, idft_h, flips = meta
        subs = [tuple(_contig(p) for p in params[_N_PARAMS * k:_N_PARAMS * (k + 1)]) for k in range(n_sub)]
        dev = subs[0][0].device
        P = sum(int(s[0].shape[0]) for s in subs)
        f32 = dict(dtype=torch.float32, device=dev)
        xyz, rot = torch.empty((P, 3), **f32), torch.empty((P, 4), **f32)
        scl, opa = torch.empty((P, 3), **f32), torch.empty((P, 1), **f32)
        feat = torch.empty((P, M, 3), **f32)
        pose = (_contig(obj_rots) if obj_rots is not None else None,
                _contig(obj_trans) if obj_trans is not None else None)
        ws = torch.empty(max(int(lib.grpg_compose_workspace_bytes(n_sub)), 256), dtype=torch.uint8, device=dev)
        desc = _descriptors(subs, is_actor, fourier, pose, idft_h, flips)
        with torch.cuda.device(dev):
            rc = lib.grpg_compose_forward(desc, n_sub, M, ws.data_ptr(), xyz.data_ptr(), rot.data_ptr(), scl.data_ptr(),
                                          opa.data_ptr(), feat.data_ptr(), _lib.current_stream_ptr(dev))
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        # Parameters and pose go through save_for_backward so that autograd's version counters catch an in-place
        # update (optimiser step, reset_opacity) between this forward and its backward, which recomputes from them.
        ctx.meta = meta
        ctx.has_pose = tuple(p is not None for p in pose)
        ctx.save_for_backward(*[t for s_ in subs for t in s_], *[p for p in pose if p is not None])
        ctx.pose_shapes = (None if obj_rots is None else obj_rots.shape, None if obj_trans is None else obj_trans.shape)
        ctx.param_shapes = [p.shape for p in params]
        return xyz, rot, scl, opa, feat

    @staticmethod
    def backward(ctx, g_xyz, g_rot, g_scl, g_opa, g_feat):
        lib = _lib.load()
        n_sub, is_actor, fourier, M, idft_h, flips = ctx.meta
        saved = ctx.saved_tensors  # raises if a saved parameter was modified in place since the forward
        subs = [tuple(saved[_N_PARAMS * k:_N_PARAMS * (k + 1)]) for k in range(n_sub)]
        rest = list(saved[_N_PARAMS * n_sub:])
        pose = tuple(rest.pop(0) if has else None for has in ctx.has_pose)
        dev = subs[0][0].device
        P = sum(int(s[0].shape[0]) for s in subs)
        f32 = dict(dtype=torch.float32, device=dev)

        def cot(g, shape):
            return torch.zeros(shape, **f32) if g is None else _contig(g)

        g_xyz, g_rot, g_scl = cot(g_xyz, (P, 3)), cot(g_rot, (P, 4)), cot(g_scl, (P, 3))
        g_opa, g_feat = cot(g_opa, (P, 1)), cot(g_feat, (P, M, 3))
        grads = [tuple(torch.empty_like(t) for t in s) for s in subs]
        views = [g for gs in grads for g in gs]
        d_rots, d_trans = torch.empty((n_sub, 4), **f32), torch.empty((n_sub, 3), **f32)
        ws = torch.empty(max(int(lib.grpg_compose_workspace_bytes(n_sub)), 256), dtype=torch.uint8, device=dev)
        desc = _descriptors(subs, is_actor, fourier, pose, idft_h, flips, grads)
        with torch.cuda.device(dev):
            rc = lib.grpg_compose_backward(desc, n_sub, M, ws.data_ptr(), g_xyz.data_ptr(), g_rot.data_ptr(),
                                           g_scl.data_ptr(), g_opa.data_ptr(), g_feat.data_ptr(), d_rots.data_ptr(),
                                           d_trans.data_ptr(), _lib.current_stream_ptr(dev))
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        n_bk = n_sub - sum(is_actor)  # actors follow the background in the table
        rs, ts = ctx.pose_shapes
        g_obj_rots = d_rots[n_bk:].reshape(rs) if rs is not None and n_sub > n_bk else None
        g_obj_trans = d_trans[n_bk:].reshape(ts) if ts is not None and n_sub > n_bk else None
        return (None, g_obj_rots, g_obj_trans, *(v.view(shape) for v, shape in zip(views, ctx.param_shapes)))


def compose_scene(background: Optional[SubModel], actors: Sequence[SubModel] = (),
                  obj_rots: Optional[torch.Tensor] = None, obj_trans: Optional[torch.Tensor] = None,
                  idft: Optional[Sequence[Sequence[float]]] = None,
                  flip_masks: Optional[Sequence[Optional[torch.Tensor]]] = None) -> ComposedScene:
    """background first (if visible), then the actors of `graph_obj_list` in order -- the reference's concatenation
    order (`graph_gaussian_range`, street_gaussian_model.py:248-262).

    obj_rots [K,4], obj_trans [K,3]: per-actor pose as `parse_camera` builds it (:265-281), BEFORE the per-Gaussian
    expand.  idft[k]: `idft_base(time_k, fourier_dim)` (a list of floats).  flip_masks[k]: bool/uint8 [n_k] or None
    (:284-293)."""
    actors = list(actors)
    K = len(actors)
    subs = ([background] if background is not None else []) + actors
    if not subs:
        raise RuntimeError("compose_scene: no visible sub-model")
    if not subs[0].xyz.is_cuda:
        raise RuntimeError("gaussianrpg_b200.scene_compose has no CPU path: parameters must be CUDA tensors")
    M = int(subs[0].features_rest.shape[1]) + 1
    is_actor = [0] * (len(subs) - K) + [1] * K
    fourier = [int(s.features_dc.shape[1]) for s in subs]
    for s in subs:
        n = s.xyz.shape[0]
        if (tuple(s.xyz.shape) != (n, 3) or tuple(s.scaling.shape) != (n, 3) or tuple(s.rotation.shape) != (n, 4)
                or s.opacity.numel() != n or tuple(s.features_rest.shape) != (n, M - 1, 3)
                or s.features_dc.shape[0] != n or s.features_dc.shape[2] != 3):
            raise RuntimeError("compose_scene: sub-model parameter shapes disagree")
    if background is not None and fourier[0] != 1:
        raise RuntimeError("compose_scene: the background has one dc row (no Fourier features)")
    if K:
        if obj_rots is None or obj_trans is None or tuple(obj_rots.shape) != (K, 4) or tuple(obj_trans.shape) != (K, 3):
            raise RuntimeError("compose_scene: obj_rots must be [K,4] and obj_trans [K,3]")
        idft_h = [list(map(float, idft[k])) if idft is not None else [1.0] + [0.0] * (fourier[len(subs) - K + k] - 1)
                  for k in range(K)]
        for k in range(K):
            if len(idft_h[k]) != fourier[len(subs) - K + k]:
                raise RuntimeError("compose_scene: idft[k] must have fourier_dim entries")
        flips = []
        for k in range(K):
            f = flip_masks[k] if flip_masks is not None else None
            if f is not None:
                if f.numel() != actors[k].xyz.shape[0]:
                    raise RuntimeError("compose_scene: flip mask size")
                f = f.to(device=actors[k].xyz.device, dtype=torch.uint8).contiguous()
            flips.append(f)
    else:
        idft_h, flips = [], []
    meta = (len(subs), is_actor, fourier, M, idft_h, flips)
    params = [t for s in subs for t in s]
    out = _ComposeScene.apply(meta, obj_rots, obj_trans, *params)
    return ComposedScene(*out)
