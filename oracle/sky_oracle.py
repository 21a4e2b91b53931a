"""CPU oracle of the sky cube-map lookup (test infrastructure only: imported by tests/ and nothing else).

Restates SkyCubeMap.forward (/root/reference/lib/models/sky_cubemap.py:77-124): ray directions as
lib/utils/graphics_utils.py:186-207 (get_rays_torch) computes them, a bilinear cube-map lookup, the foreground mask
((1 - acc) > 1e-3, :86-87), the fill colour outside the mask (:104-115), clamp to [0, 1] and CHW output (:117).

PARITY UNPINNED.  The lookup itself is nvdiffrast's `dr.texture(..., filter_mode='linear', boundary_mode='cube')`
(sky_cubemap.py:7,98,111), a third-party package that is not vendored in /root/reference and not installed in this
image, so no golden vectors of the reference's own output can be produced here.  What is restated is the published
cube mapping: major-axis face selection in the OpenGL face order and orientation (+x,-x,+y,-y,+z,-z), texel centres at
(i + 0.5) / res, bilinear weights, seamless filtering across face edges (a tap that falls off a face comes from the
adjacent face's texel at the same position along the shared edge) and renormalisation of the three remaining taps at a
cube corner.  This file is written independently of the CUDA kernel (float64, whole-array numpy, a different
derivation of the edge taps: by 3D rotation of the unfolded neighbour) so that the two check each other.
"""
from __future__ import annotations

import numpy as np

# face -> (major axis, sign, (axis of u, sign of u), (axis of v, sign of v)); u = sign_u * p[axis_u] / |major| etc.
FACES = [(0, +1, (2, -1), (1, -1)), (0, -1, (2, +1), (1, -1)), (1, +1, (0, +1), (2, +1)),
         (1, -1, (0, +1), (2, -1)), (2, +1, (0, +1), (1, -1)), (2, -1, (0, -1), (1, -1))]


def ray_directions(H, W, K, R, T, jitter=None):
    """graphics_utils.py:186-207 in float64; jitter [2,H,W] = (perturb_i, perturb_j) or None for pixel centres."""
    K, R, T = np.asarray(K, np.float64), np.asarray(R, np.float64), np.asarray(T, np.float64).reshape(3)
    i, j = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    ji, jj = (0.5, 0.5) if jitter is None else (np.asarray(jitter[0], np.float64), np.asarray(jitter[1], np.float64))
    xy1 = np.stack([i + ji, j + jj, np.ones_like(i)], 2)
    rays_o = -R.T @ T
    pixel_camera = xy1 @ np.linalg.inv(K).T
    pixel_world = (pixel_camera - T) @ R
    d = pixel_world - rays_o
    return d / np.linalg.norm(d, axis=2, keepdims=True)


def _face_uv(d):
    """d [...,3] -> (face, u, v) with u, v in [0, 1]"""
    ad = np.abs(d)
    face = np.where(ad[..., 2] > np.maximum(ad[..., 0], ad[..., 1]), 4, np.where(ad[..., 1] > ad[..., 0], 2, 0))
    major = np.take_along_axis(d, (face // 2)[..., None], -1)[..., 0]
    face = face + (major < 0)
    u = np.empty(d.shape[:-1]); v = np.empty(d.shape[:-1])
    for f, (ax, sg, (au, su), (av, sv)) in enumerate(FACES):
        m = face == f
        u[m] = su * d[..., au][m] / (2 * np.abs(d[..., ax][m])) + 0.5
        v[m] = sv * d[..., av][m] / (2 * np.abs(d[..., ax][m])) + 0.5
    return face, np.clip(u, 0, 1), np.clip(v, 0, 1)


def _texel(face, i, j, res):
    """linear texel index of (face, i, j) with i or j possibly one step off the face; -1 at a corner"""
    face, i, j = np.broadcast_arrays(face, i, j)
    out = np.full(face.shape, -1, np.int64)
    oi, oj = (i < 0) | (i >= res), (j < 0) | (j >= res)
    ins = ~oi & ~oj
    out[ins] = (face[ins] * res + j[ins]) * res + i[ins]
    edge = oi ^ oj
    if edge.any():
        f, ii, jj = face[edge], i[edge].astype(np.float64), j[edge].astype(np.float64)
        a, b = (2 * ii + 1) / res - 1, (2 * jj + 1) / res - 1  # face coordinates of the texel centre in [-1-1/res, 1+1/res]
        p = np.zeros(f.shape + (3,))
        for k, (ax, sg, (au, su), (av, sv)) in enumerate(FACES):
            m = f == k
            p[m, ax] = sg
            p[m, au] = su * a[m]
            p[m, av] = sv * b[m]
        # The neighbour lies on the plane of this face, 1/res beyond the edge.  Rotate it by 90 degrees about the edge
        # onto the adjacent face: its distance beyond the edge becomes its depth below the edge on that face.
        major = (f // 2)
        q = p.copy()
        for n in range(p.shape[0]):
            A = major[n]
            B = [ax for ax in range(3) if ax != A and abs(p[n, ax]) > 1][0]
            over = abs(p[n, B]) - 1.0
            q[n, B] = np.sign(p[n, B])
            q[n, A] = p[n, A] * (1.0 - over)
        f2, u2, v2 = _face_uv(q)
        i2 = np.minimum(res - 1, np.floor(u2 * res).astype(np.int64))
        j2 = np.minimum(res - 1, np.floor(v2 * res).astype(np.int64))
        out[edge] = (f2 * res + j2) * res + i2
    return out


def taps(d, res):
    """four (texel index, weight) pairs per direction: [...,4] int64 and [...,4] float64 (missing taps: weight 0)"""
    face, u, v = _face_uv(d)
    fu, fv = u * res - 0.5, v * res - 0.5
    i0, j0 = np.floor(fu).astype(np.int64), np.floor(fv).astype(np.int64)
    wu, wv = fu - i0, fv - j0
    idx = np.stack([_texel(face, i0, j0, res), _texel(face, i0 + 1, j0, res), _texel(face, i0, j0 + 1, res),
                    _texel(face, i0 + 1, j0 + 1, res)], -1)
    w = np.stack([(1 - wu) * (1 - wv), wu * (1 - wv), (1 - wu) * wv, wu * wv], -1)
    w = np.where(idx < 0, 0.0, w)
    return idx, w / w.sum(-1, keepdims=True)


def sky_forward(cubemap, H, W, K, R, T, acc=None, mask=None, fill=0.0, jitter=None):
    """-> (sky [3,H,W] clamped, unclamped [H,W,3], mask [H,W])"""
    cube = np.asarray(cubemap, np.float64)
    res = cube.shape[1]
    d = ray_directions(H, W, K, R, T, jitter)
    if mask is None:
        mask = np.ones((H, W), bool) if acc is None else (1.0 - np.asarray(acc, np.float32).reshape(H, W)) > np.float32(1e-3)
    idx, w = taps(d, res)
    flat = cube.reshape(-1, 3)
    col = (flat[np.maximum(idx, 0)] * w[..., None]).sum(-2)
    col = np.where(mask[..., None], col, fill)
    return np.clip(col, 0, 1).transpose(2, 0, 1), col, mask


def sky_backward(cubemap, H, W, K, R, T, dL_dsky, acc=None, mask=None, fill=0.0, jitter=None):
    """gradient of sum(sky * dL_dsky) with respect to the cube map: [6,res,res,3]"""
    cube = np.asarray(cubemap, np.float64)
    res = cube.shape[1]
    _, col, mask = sky_forward(cubemap, H, W, K, R, T, acc, mask, fill, jitter)
    d = ray_directions(H, W, K, R, T, jitter)
    idx, w = taps(d, res)
    g = np.asarray(dL_dsky, np.float64).transpose(1, 2, 0) * ((col >= 0) & (col <= 1)) * mask[..., None]
    out = np.zeros((6 * res * res, 3))
    for k in range(4):
        ok = idx[..., k] >= 0
        np.add.at(out, idx[..., k][ok], (w[..., k][ok])[:, None] * g[ok])
    return out.reshape(6, res, res, 3)
