"""CPU oracle (TEST INFRASTRUCTURE, never on the product path) for the scene-graph compose op, SURVEY 8f rank 1.

numpy float64 restatement of what `StreetGaussianModel` hands to the rasterizer after `parse_camera`
(/root/reference/lib/models/street_gaussian_model.py): get_xyz :341-367, get_rotation :314-338, get_scaling :296-312,
get_opacity :438-453, get_features :370-384, with the sub-model getters they call -- exp / sigmoid / F.normalize
(gaussian_model.py:214-251) and the actors' Fourier dc (gaussian_model_actor.py:73-82) -- and the helpers
quaternion_raw_multiply (lib/utils/general_utils.py:220-238) and quaternion_to_matrix (:125-146).  The backward is
derived by hand (no autograd).  Pinned by tests/golden/compose_*.npz, which hold outputs and autograd gradients of the
reference's own getters (tests/golden/make_compose_golden.py).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12  # torch.nn.functional.normalize default
FLIP_QUAT = np.array([0.0, 0.0, 1.0, 0.0])  # matrix_to_quaternion(diag(-1, 1, -1)), street_gaussian_model.py:59-61
FLIP_AXIS = 1


def qmul(a, b):
    """quaternion_raw_multiply, general_utils.py:232-238 (real part first)."""
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def qmul_bwd(a, b, g):
    """cotangents of a and b for o = qmul(a, b)."""
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    gw, gx, gy, gz = g[..., 0], g[..., 1], g[..., 2], g[..., 3]
    da = np.stack([bw * gw + bx * gx + by * gy + bz * gz, -bx * gw + bw * gx - bz * gy + by * gz,
                   -by * gw + bz * gx + bw * gy - bx * gz, -bz * gw - by * gx + bx * gy + bw * gz], -1)
    db = np.stack([aw * gw + ax * gx + ay * gy + az * gz, -ax * gw + aw * gx + az * gy - ay * gz,
                   -ay * gw - az * gx + aw * gy + ax * gz, -az * gw + ay * gx - ax * gy + aw * gz], -1)
    return da, db


def normalize(v):
    n = np.maximum(np.sqrt((v * v).sum(-1, keepdims=True)), EPS)
    return v / n


def normalize_bwd(v, g):
    """F.normalize backward: below eps the denominator is the constant eps."""
    nrm = np.sqrt((v * v).sum(-1, keepdims=True))
    n = np.maximum(nrm, EPS)
    y = v / n
    return np.where(nrm > EPS, (g - y * (y * g).sum(-1, keepdims=True)) / n, g / EPS)


def quat_to_matrix(q):
    """quaternion_to_matrix, general_utils.py:125-146 (normalises first, no eps)."""
    qn = q / np.sqrt((q * q).sum())
    r, x, y, z = qn
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                     [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                     [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])


def quat_to_matrix_bwd(q, G):
    nrm = np.sqrt((q * q).sum())
    r, x, y, z = q / nrm
    d = 2.0 * np.array([
        -z * G[0, 1] + y * G[0, 2] + z * G[1, 0] - x * G[1, 2] - y * G[2, 0] + x * G[2, 1],
        y * G[0, 1] + z * G[0, 2] + y * G[1, 0] - 2 * x * G[1, 1] - r * G[1, 2] + z * G[2, 0] + r * G[2, 1] - 2 * x * G[2, 2],
        -2 * y * G[0, 0] + x * G[0, 1] + r * G[0, 2] + x * G[1, 0] + z * G[1, 2] - r * G[2, 0] + z * G[2, 1] - 2 * y * G[2, 2],
        -2 * z * G[0, 0] - r * G[0, 1] + x * G[0, 2] + r * G[1, 0] - 2 * z * G[1, 1] + y * G[1, 2] + x * G[2, 0] + y * G[2, 1]])
    qn = q / nrm
    return (d - qn * (qn * d).sum()) / nrm


def _f64(sub):
    return {k: np.asarray(v, dtype=np.float64) for k, v in sub.items()}


def compose(bkgd, actors, obj_rots, obj_trans, idfts, flips):
    """-> dict(xyz, rotation, scaling, opacity, features), background first then the actors in order."""
    X, Q, S, O, Fe = [], [], [], [], []
    if bkgd is not None:
        b = _f64(bkgd)
        X.append(b["xyz"]); Q.append(normalize(b["rotation"])); S.append(np.exp(b["scaling"]))
        O.append(1.0 / (1.0 + np.exp(-b["opacity"])))
        Fe.append(np.concatenate([b["features_dc"], b["features_rest"]], 1))  # gaussian_model.py:237-240
    for k, sub in enumerate(actors):
        a = _f64(sub)
        flip = np.asarray(flips[k], dtype=bool)
        qo = np.asarray(obj_rots[k], dtype=np.float64)
        xl = a["xyz"].copy()
        xl[flip, FLIP_AXIS] *= -1                                           # street_gaussian_model.py:360
        X.append(xl @ quat_to_matrix(qo).T + np.asarray(obj_trans[k], dtype=np.float64))  # :361-362
        ql = normalize(a["rotation"])                                       # gaussian_model.py:229-230
        ql[flip] = qmul(FLIP_QUAT, ql[flip])                                # street_gaussian_model.py:332
        Q.append(normalize(qmul(qo, ql)))                                   # :333-334
        S.append(np.exp(a["scaling"])); O.append(1.0 / (1.0 + np.exp(-a["opacity"])))
        dc = (a["features_dc"] * np.asarray(idfts[k], dtype=np.float64)[None, :, None]).sum(1, keepdims=True)
        Fe.append(np.concatenate([dc, a["features_rest"]], 1))             # gaussian_model_actor.py:77-82
    cat = lambda l: np.concatenate(l, 0)
    return dict(xyz=cat(X), rotation=cat(Q), scaling=cat(S), opacity=cat(O), features=cat(Fe))


def compose_bwd(bkgd, actors, obj_rots, obj_trans, idfts, flips, g):
    """Cotangents g[xyz|rotation|scaling|opacity|features] -> (per-sub-model parameter grads, grad_obj_rots,
    grad_obj_trans).  Sub-model order: background (if any) then actors."""
    g = _f64(g)
    grads, off = [], 0
    K = len(actors)
    g_rots, g_trans = np.zeros((K, 4)), np.zeros((K, 3))

    def common(p, sl):
        sig = 1.0 / (1.0 + np.exp(-p["opacity"]))
        return dict(scaling=g["scaling"][sl] * np.exp(p["scaling"]), opacity=g["opacity"][sl] * sig * (1 - sig),
                    features_rest=g["features"][sl][:, 1:])

    if bkgd is not None:
        b = _f64(bkgd)
        n = b["xyz"].shape[0]
        sl = slice(0, n)
        d = common(b, sl)
        d["xyz"] = g["xyz"][sl]
        d["rotation"] = normalize_bwd(b["rotation"], g["rotation"][sl])
        d["features_dc"] = g["features"][sl][:, :1]
        grads.append(d)
        off = n
    for k, sub in enumerate(actors):
        a = _f64(sub)
        n = a["xyz"].shape[0]
        sl = slice(off, off + n)
        off += n
        flip = np.asarray(flips[k], dtype=bool)
        qo = np.asarray(obj_rots[k], dtype=np.float64)
        d = common(a, sl)
        R = quat_to_matrix(qo)
        xl = a["xyz"].copy()
        xl[flip, FLIP_AXIS] *= -1
        gx = g["xyz"][sl]
        dx = gx @ R
        dx[flip, FLIP_AXIS] *= -1
        d["xyz"] = dx
        g_trans[k] = gx.sum(0)
        g_rots[k] = quat_to_matrix_bwd(qo, gx.T @ xl)
        ql0 = normalize(a["rotation"])
        ql = ql0.copy()
        ql[flip] = qmul(FLIP_QUAT, ql[flip])
        gq = normalize_bwd(qmul(qo, ql), g["rotation"][sl])
        da, dql = qmul_bwd(np.broadcast_to(qo, ql.shape), ql, gq)
        g_rots[k] += da.sum(0)
        _, dfl = qmul_bwd(np.broadcast_to(FLIP_QUAT, ql.shape), ql0, dql)
        dql = np.where(flip[:, None], dfl, dql)
        d["rotation"] = normalize_bwd(a["rotation"], dql)
        d["features_dc"] = g["features"][sl][:, :1] * np.asarray(idfts[k], dtype=np.float64)[None, :, None]
        grads.append(d)
    return grads, g_rots, g_trans
