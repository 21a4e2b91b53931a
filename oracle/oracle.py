"""numpy front-end of the CPU oracle (oracle/grpg_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "grpg_oracle.c"
LIB = HERE / "libgrpg_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    base = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared", str(SRC), "-o", str(LIB), "-lm"]
    # -mfma makes fmaf() a single instruction; without it glibc's (correct, slower) fmaf is used
    for extra in (["-mfma"], []):
        r = subprocess.run(base[:1] + extra + base[1:], capture_output=True, text=True)
        if r.returncode == 0:
            # a binary built with -mfma must also run here
            try:
                C.CDLL(str(LIB)).oracle_version()
                return LIB
            except OSError:
                continue
    raise RuntimeError(f"gcc failed to build the oracle:\n{r.stderr}")


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def preprocess(means3D, opacities, view, proj, campos, W, H, tan_fovx, tan_fovy, *, shs=None, sh_degree=0,
               colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0):
    lib = load()
    means3D, opacities = _f(means3D), _f(opacities).reshape(-1)
    P = means3D.shape[0]
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(_f, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    M = 0 if shs is None else shs.shape[1]
    view, proj, campos = _f(view).reshape(-1), _f(proj).reshape(-1), _f(campos).reshape(-1)
    out = dict(
        radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
        cov3D=np.zeros((P, 6), np.float32) if cov3D_precomp is None else cov3D_precomp.copy(),
        rgb=np.zeros((P, 3), np.float32) if colors_precomp is None else colors_precomp.copy(),
        conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
        clamped=np.zeros((P, 3), np.uint8), rects=np.zeros((P, 4), np.uint32))
    lib.oracle_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales), C.c_float(scale_modifier), _p(rotations),
        _p(opacities), _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(view), _p(proj), _p(campos), C.c_int(W),
        C.c_int(H), C.c_float(tan_fovx), C.c_float(tan_fovy), _p(out["radii"]), _p(out["means2D"]), _p(out["depths"]),
        _p(out["cov3D"]), _p(out["rgb"]), _p(out["conic_opacity"]), _p(out["tiles_touched"]), _p(out["clamped"]),
        _p(out["rects"]))
    return out


def binning(pre, W, H):
    lib = load()
    lib.oracle_binning.restype = C.c_longlong
    P = pre["radii"].shape[0]
    cap = int(pre["tiles_touched"].astype(np.int64).sum())
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    keys = np.zeros(max(cap, 1), np.uint64)
    point_list = np.zeros(max(cap, 1), np.uint32)
    ranges = np.zeros((tiles, 2), np.uint32)
    R = lib.oracle_binning(C.c_int(P), C.c_int(W), C.c_int(H), _p(pre["radii"]), _p(pre["depths"]), _p(pre["rects"]),
                           _p(keys), _p(point_list), _p(ranges))
    assert R == cap
    return dict(R=int(R), keys=keys[:R], point_list=point_list[:R], ranges=ranges)


def blend_forward(pre, binned, W, H, bg, semantics=None, tile_rows=None):
    lib = load()
    S = 0 if semantics is None else semantics.shape[1]
    semantics = _f(semantics)
    bg = _f(bg).reshape(-1)
    gy = (H + 15) // 16
    y0, y1 = (0, gy) if tile_rows is None else tile_rows
    out = dict(color=np.zeros((3, H, W), np.float32), depth=np.zeros((1, H, W), np.float32),
               alpha=np.zeros((1, H, W), np.float32), semantic=np.zeros((S, H, W), np.float32),
               n_contrib=np.zeros((H, W), np.uint32))
    lib.oracle_blend_forward(
        C.c_int(W), C.c_int(H), C.c_int(S), C.c_int(y0), C.c_int(y1), _p(binned["ranges"]), _p(binned["point_list"]),
        _p(pre["means2D"]), _p(pre["rgb"]), _p(pre["depths"]), _p(semantics), _p(pre["conic_opacity"]), _p(bg),
        _p(out["color"]), _p(out["depth"]), _p(out["alpha"]), _p(out["semantic"]), _p(out["n_contrib"]))
    return out


def forward(means3D, opacities, view, proj, campos, W, H, tan_fovx, tan_fovy, bg, *, semantics=None, **kw):
    pre = preprocess(means3D, opacities, view, proj, campos, W, H, tan_fovx, tan_fovy, **kw)
    binned = binning(pre, W, H)
    img = blend_forward(pre, binned, W, H, bg, semantics)
    return pre, binned, img


def backward(means3D, view, proj, campos, W, H, tan_fovx, tan_fovy, bg, pre, binned, img, dL_dcolor, dL_ddepth,
             dL_dalpha, dL_dsemantic=None, *, semantics=None, shs=None, sh_degree=0, scales=None, rotations=None,
             scale_modifier=1.0):
    lib = load()
    means3D = _f(means3D)
    P = means3D.shape[0]
    S = 0 if semantics is None else semantics.shape[1]
    semantics, shs, scales, rotations = map(_f, (semantics, shs, scales, rotations))
    M = 0 if shs is None else shs.shape[1]
    bg, view, proj, campos = _f(bg).reshape(-1), _f(view).reshape(-1), _f(proj).reshape(-1), _f(campos).reshape(-1)
    dL_dcolor, dL_ddepth, dL_dalpha = _f(dL_dcolor), _f(dL_ddepth), _f(dL_dalpha)
    dL_dsemantic = _f(dL_dsemantic) if S > 0 else np.zeros((0, H, W), np.float32)
    acc = dict(mean2D=np.zeros((P, 3)), conic=np.zeros((P, 4)), opacity=np.zeros(P), colors=np.zeros((P, 3)),
               depths=np.zeros(P), semantics=np.zeros((P, S)))
    lib.oracle_blend_backward(
        C.c_int(W), C.c_int(H), C.c_int(S), _p(binned["ranges"]), _p(binned["point_list"]), _p(bg), _p(pre["means2D"]),
        _p(pre["conic_opacity"]), _p(pre["rgb"]), _p(pre["depths"]), _p(semantics), _p(img["alpha"]),
        _p(img["n_contrib"]), _p(dL_dcolor), _p(dL_ddepth), _p(dL_dalpha), _p(dL_dsemantic), _p(acc["mean2D"]),
        _p(acc["conic"]), _p(acc["opacity"]), _p(acc["colors"]), _p(acc["depths"]), _p(acc["semantics"]))
    g = {k: v.astype(np.float32) for k, v in acc.items()}
    out = dict(dL_dmeans2D=g["mean2D"], dL_dconic=g["conic"], dL_dopacity=g["opacity"].reshape(P, 1),
               dL_dcolors=g["colors"], dL_ddepths=g["depths"].reshape(P, 1), dL_dsemantics=g["semantics"],
               dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
               dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
               dL_drotations=np.zeros((P, 4), np.float32))
    lib.oracle_preprocess_backward(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(pre["radii"]), _p(shs), _p(pre["clamped"]),
        _p(scales), _p(rotations), C.c_float(scale_modifier), _p(pre["cov3D"]), _p(view), _p(proj), C.c_int(W),
        C.c_int(H), C.c_float(tan_fovx), C.c_float(tan_fovy), _p(campos), _p(out["dL_dmeans2D"]), _p(out["dL_dconic"]),
        _p(out["dL_dcolors"]), _p(g["depths"]), _p(out["dL_dmeans3D"]), _p(out["dL_dcov3D"]), _p(out["dL_dsh"]),
        _p(out["dL_dscales"]), _p(out["dL_drotations"]))
    return out


def knn_mean_dist2(points):
    """simple-knn distCUDA2 by brute force (oracle_knn_mean_dist2): float32 [P]."""
    lib = load()
    pts = _f(points).reshape(-1, 3)
    out = np.empty(pts.shape[0], dtype=np.float32)
    lib.oracle_knn_mean_dist2(C.c_int(pts.shape[0]), _p(pts), _p(out))
    return out
