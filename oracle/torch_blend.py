"""Pure-PyTorch per-pixel CPU alpha-blend of a Gaussian scene.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

BASELINE.json asks for "a pure-PyTorch per-pixel CPU alpha-blend of the same scene timed on the box's
own host cores" next to the GPU numbers, and SURVEY 8(c) lists what a naive blend gets wrong; this
module honours those rules (tile-rect membership, stable (tile, depth bits, index) order, the
T*(1-alpha) < 1e-4 stop rule, un-normalised depth, alpha = sum(alpha_i T_i)) and is differentiable
with torch.autograd so it doubles as an independent check of the C oracle's gradients:
  * straight-through min(0.99, .)                      (backward.cu:543,618)
  * clamped t.x / t.y frozen, not fed back into t.z     (backward.cu:175-176,262-264)
  * dL/dmeans2D reported in NDC units (x 0.5 W, 0.5 H)   (backward.cu:501-502,625-626)
Only tests/ and bench.py's cpu_baseline leg import it.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def _sh_color(deg, sh, dirs):
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                   + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                       + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11]
                       + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                       + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                       + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res + 0.5


def preprocess(means3D, opacities, view, proj, campos, W, H, tanx, tany, *, shs=None, sh_degree=0,
               colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
               means2D_offset: Optional[torch.Tensor] = None):
    """Vectorised per-Gaussian stage.  Returns a dict of per-Gaussian tensors (all P rows) + `visible`."""
    dt = means3D.dtype
    P = means3D.shape[0]
    fx, fy = W / (2.0 * tanx), H / (2.0 * tany)
    ones = torch.ones(P, 1, dtype=dt)
    hom = torch.cat([means3D, ones], 1)
    pv = hom @ view          # the matrices are passed transposed: row-vector convention
    ph = hom @ proj
    tz = pv[:, 2]
    visible = tz > 0.2
    p_w = 1.0 / (ph[:, 3] + 1e-7)
    ndc = ph[:, :2] * p_w[:, None]
    if cov3D_precomp is not None:
        c = cov3D_precomp
        Sigma = torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], 1), torch.stack([c[:, 1], c[:, 3], c[:, 4]], 1),
                             torch.stack([c[:, 2], c[:, 4], c[:, 5]], 1)], 1)
    else:
        r, x, y, z = rotations.unbind(1)  # used un-normalised (forward.cu:127)
        Rm = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
                          torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
                          torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)
        s = scale_modifier * scales
        Sigma = Rm @ torch.diag_embed(s * s) @ Rm.transpose(1, 2)
    limx, limy = 1.3 * tanx, 1.3 * tany
    safe_tz = torch.where(visible, tz, torch.ones_like(tz))
    txtz, tytz = pv[:, 0] / safe_tz, pv[:, 1] / safe_tz
    cl_x, cl_y = (txtz < -limx) | (txtz > limx), (tytz < -limy) | (tytz > limy)
    tx = torch.where(cl_x, (txtz.clamp(-limx, limx) * safe_tz).detach(), pv[:, 0])
    ty = torch.where(cl_y, (tytz.clamp(-limy, limy) * safe_tz).detach(), pv[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / safe_tz, zero, -(fx * tx) / (safe_tz * safe_tz)], 1),
                     torch.stack([zero, fy / safe_tz, -(fy * ty) / (safe_tz * safe_tz)], 1)], 1)  # [P,2,3]
    Wm = view[:3, :3].t()  # rotation part, so that t = Wm @ p + trans
    JW = J @ Wm
    cov2 = JW @ Sigma @ JW.transpose(1, 2)
    a, b, c = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    visible = visible & (det != 0)
    det_s = torch.where(det != 0, det, torch.ones_like(det))
    conic = torch.stack([c / det_s, -b / det_s, a / det_s], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    pix = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    if means2D_offset is not None:
        pix = pix + means2D_offset
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pd = pix.detach()
    x0 = torch.clamp(torch.trunc((pd[:, 0] - radius) / 16), 0, gx).long()
    y0 = torch.clamp(torch.trunc((pd[:, 1] - radius) / 16), 0, gy).long()
    x1 = torch.clamp(torch.trunc((pd[:, 0] + radius + 15) / 16), 0, gx).long()
    y1 = torch.clamp(torch.trunc((pd[:, 1] + radius + 15) / 16), 0, gy).long()
    visible = visible & ((x1 - x0) * (y1 - y0) > 0)
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - campos[None, :]
        rgb = torch.clamp(_sh_color(sh_degree, shs, d / d.norm(dim=1, keepdim=True)), min=0.0)
    return dict(visible=visible, depth=tz, pix=pix, conic=conic, opacity=opacities.reshape(-1), rgb=rgb,
                radius=radius.long() * visible, rect=(x0, y0, x1, y1))


def blend(pre, W, H, bg, semantics=None, tile_subset=None):
    """Per-tile sequential blend, vectorised over the 256 pixels of a tile.  Returns dict of images."""
    dt = pre["pix"].dtype
    gx, gy = (W + 15) // 16, (H + 15) // 16
    x0, y0, x1, y1 = pre["rect"]
    vis_idx = torch.nonzero(pre["visible"]).flatten()
    # instance list in the reference order: stable sort of (tile, depth float bits) over row-major emission
    depth_bits = pre["depth"].detach().float().view(torch.int32).long()
    tiles, gids = [], []
    for g in vis_idx.tolist():
        ty = torch.arange(int(y0[g]), int(y1[g]))
        tx = torch.arange(int(x0[g]), int(x1[g]))
        t = (ty[:, None] * gx + tx[None, :]).flatten()
        tiles.append(t)
        gids.append(torch.full_like(t, g))
    S = 0 if semantics is None else semantics.shape[1]
    color = torch.zeros(3, H, W, dtype=dt)
    depth = torch.zeros(1, H, W, dtype=dt)
    alpha = torch.zeros(1, H, W, dtype=dt)
    sem = torch.zeros(S, H, W, dtype=dt)
    n_contrib = torch.zeros(H, W, dtype=torch.int64)
    if not tiles:
        color = color + bg[:, None, None]
        return dict(color=color, depth=depth, alpha=alpha, semantic=sem, n_contrib=n_contrib, R=0)
    tiles = torch.cat(tiles)
    gids = torch.cat(gids)
    key = (tiles << 32) | depth_bits[gids]
    order = torch.sort(key, stable=True).indices
    tiles, gids = tiles[order], gids[order]
    out_c, out_d, out_a, out_s = [], [], [], []
    tile_ids = range(gx * gy) if tile_subset is None else tile_subset
    bounds = torch.searchsorted(tiles, torch.arange(gx * gy + 1))
    for t in tile_ids:
        ty, tx = divmod(t, gx)
        py = torch.arange(ty * 16, min(ty * 16 + 16, H))
        px = torch.arange(tx * 16, min(tx * 16 + 16, W))
        PX = px[None, :].expand(len(py), len(px)).to(dt)
        PY = py[:, None].expand(len(py), len(px)).to(dt)
        T = torch.ones_like(PX)
        done = torch.zeros_like(PX, dtype=torch.bool)
        C = torch.zeros(3, *PX.shape, dtype=dt)
        D = torch.zeros_like(PX)
        Wt = torch.zeros_like(PX)
        Sm = torch.zeros(S, *PX.shape, dtype=dt)
        last = torch.zeros_like(PX, dtype=torch.int64)
        lo, hi = int(bounds[t]), int(bounds[t + 1])
        for n, k in enumerate(range(lo, hi), start=1):
            g = int(gids[k])
            dx, dy = pre["pix"][g, 0] - PX, pre["pix"][g, 1] - PY
            cA, cB, cC = pre["conic"][g]
            power = -0.5 * (cA * dx * dx + cC * dy * dy) - cB * dx * dy
            a_raw = pre["opacity"][g] * torch.exp(power)
            a = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()  # straight-through clamp
            ok = (~done) & (power <= 0) & (a >= 1.0 / 255.0)
            test_T = T * (1 - a)
            stop = ok & (test_T < 1e-4)
            done = done | stop
            ok = ok & ~stop
            w = torch.where(ok, a * T, torch.zeros_like(T))
            C = C + pre["rgb"][g][:, None, None] * w
            D = D + pre["depth"][g] * w
            Wt = Wt + w
            if S:
                Sm = Sm + semantics[g][:, None, None] * w
            T = torch.where(ok, test_T, T)
            last = torch.where(ok, torch.full_like(last, n), last)
            if bool(done.all()):
                break
        ys, xs = slice(int(py[0]), int(py[-1]) + 1), slice(int(px[0]), int(px[-1]) + 1)
        out_c.append((ys, xs, C + T * bg[:, None, None]))
        out_d.append((ys, xs, D))
        out_a.append((ys, xs, Wt))
        out_s.append((ys, xs, Sm))
        n_contrib[ys, xs] = last
    # assemble without in-place ops on graph tensors
    def assemble(parts, ch):
        rows = []
        img = torch.zeros(ch, H, W, dtype=dt)
        mask = torch.zeros(1, H, W, dtype=torch.bool)
        for ys, xs, v in parts:
            pad = torch.zeros(ch, H, W, dtype=dt)
            pad[:, ys, xs] = v.reshape(ch, v.shape[-2], v.shape[-1]) if v.dim() == 3 else v[None]
            rows.append(pad)
            mask[:, ys, xs] = True
        return (torch.stack(rows).sum(0) if rows else img), mask
    color, mask = assemble(out_c, 3)
    if tile_subset is None:
        pass
    depth, _ = assemble(out_d, 1)
    alpha, _ = assemble(out_a, 1)
    sem = assemble(out_s, S)[0] if S else sem
    return dict(color=color, depth=depth, alpha=alpha, semantic=sem, n_contrib=n_contrib, R=int(tiles.numel()),
                mask=mask)


def render(sc, dtype=torch.float64, requires_grad=False, tile_subset=None):
    """Render a gaussianrpg_b200.synthetic.Scene on the CPU; returns (images, leaves)."""
    cvt = lambda t: None if t is None else t.detach().to(dtype).clone().requires_grad_(requires_grad)  # noqa: E731
    leaves = dict(means3D=cvt(sc.means3D), opacities=cvt(sc.opacities), shs=cvt(sc.shs), scales=cvt(sc.scales),
                  rotations=cvt(sc.rotations), colors_precomp=cvt(sc.colors_precomp),
                  cov3D_precomp=cvt(sc.cov3D_precomp), semantics=cvt(sc.semantics))
    P = sc.means3D.shape[0]
    leaves["means2D"] = torch.zeros(P, 2, dtype=dtype, requires_grad=requires_grad)
    pre = preprocess(leaves["means3D"], leaves["opacities"], sc.viewmatrix.to(dtype), sc.projmatrix.to(dtype),
                     sc.campos.to(dtype), sc.width, sc.height, sc.tanfovx, sc.tanfovy, shs=leaves["shs"],
                     sh_degree=sc.sh_degree, colors_precomp=leaves["colors_precomp"], scales=leaves["scales"],
                     rotations=leaves["rotations"], cov3D_precomp=leaves["cov3D_precomp"],
                     scale_modifier=sc.scale_modifier, means2D_offset=leaves["means2D"])
    img = blend(pre, sc.width, sc.height, sc.bg.to(dtype), leaves["semantics"], tile_subset)
    return img, leaves, pre


def time_tiles(pre_np, binned, sc, tile_ids, dtype=torch.float32, budget_s=20.0):
    """Time the per-pixel blend (forward + autograd backward) of the given tiles, starting from the C oracle's
    per-Gaussian outputs and instance order; stops once `budget_s` seconds are spent.
    Returns (seconds, instances_processed, tiles_processed).  (bench.py cpu_baseline leg.)"""
    import time as _time
    W, H = sc.width, sc.height
    gx = (W + 15) // 16
    pix = torch.from_numpy(pre_np["means2D"]).to(dtype)
    conic = torch.from_numpy(pre_np["conic_opacity"][:, :3].copy()).to(dtype)
    opac = torch.from_numpy(pre_np["conic_opacity"][:, 3].copy()).to(dtype)
    rgb = torch.from_numpy(pre_np["rgb"]).to(dtype)
    dep = torch.from_numpy(pre_np["depths"]).to(dtype)
    bg = sc.bg.to(dtype)
    plist = torch.from_numpy(binned["point_list"].astype("int64"))
    t0 = _time.time()
    n_inst = n_tiles = 0
    for t in tile_ids:
        if _time.time() - t0 > budget_s and n_tiles >= 2:
            break
        lo, hi = int(binned["ranges"][t, 0]), int(binned["ranges"][t, 1])
        n_inst += hi - lo
        n_tiles += 1
        ids = plist[lo:hi]
        lp = pix[ids].clone().requires_grad_(True)
        lc = conic[ids].clone().requires_grad_(True)
        lo_ = opac[ids].clone().requires_grad_(True)
        lr = rgb[ids].clone().requires_grad_(True)
        ld = dep[ids].clone().requires_grad_(True)
        ty, tx = divmod(t, gx)
        py = torch.arange(ty * 16, min(ty * 16 + 16, H))
        px = torch.arange(tx * 16, min(tx * 16 + 16, W))
        PX = px[None, :].expand(len(py), len(px)).to(dtype)
        PY = py[:, None].expand(len(py), len(px)).to(dtype)
        T = torch.ones_like(PX)
        done = torch.zeros_like(PX, dtype=torch.bool)
        C = torch.zeros(3, *PX.shape, dtype=dtype)
        D = torch.zeros_like(PX)
        Wt = torch.zeros_like(PX)
        for k in range(hi - lo):
            dx, dy = lp[k, 0] - PX, lp[k, 1] - PY
            power = -0.5 * (lc[k, 0] * dx * dx + lc[k, 2] * dy * dy) - lc[k, 1] * dx * dy
            a_raw = lo_[k] * torch.exp(power)
            a = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()
            ok = (~done) & (power <= 0) & (a >= 1.0 / 255.0)
            test_T = T * (1 - a)
            stop = ok & (test_T < 1e-4)
            done = done | stop
            ok = ok & ~stop
            w = torch.where(ok, a * T, torch.zeros_like(T))
            C = C + lr[k][:, None, None] * w
            D = D + ld[k] * w
            Wt = Wt + w
            T = torch.where(ok, test_T, T)
            if bool(done.all()):
                break
        loss = (C + T * bg[:, None, None]).sum() + 0.01 * D.sum() + 0.1 * Wt.sum()
        if loss.requires_grad:
            loss.backward()
    return _time.time() - t0, n_inst, n_tiles
