"""The backward blend kernels do not carry the reference's per-channel "last contributor" recurrences
(backward.cu:560-604); they carry their gradient-weighted sum:

    reference:  acc_u <- a_last u_last + (1 - a_last) acc_u   for every channel u;   dL/da_k = sum_u (u_k - acc_u) g_u
    kernels:    s_k = sum_u u_k g_u;   dL/da_k = s_k - A;   A <- a_k s_k + (1 - a_k) A

This test replays both in float32 (numpy scalars, the kernels' operation order) on pixels of the golden scenes --
including the all-cancellation `adversarial` one -- and checks that dL/da agrees to far better than the 1e-3 gradient
bar, per contribution and accumulated per Gaussian."""
import numpy as np
import pytest

import cases
from oracle import oracle

F = np.float32


def _pixel_lists(sc, n_pixels, seed):
    """(alpha_k, channels u_k [5]) back-to-front for a sample of pixels, exactly as the blend decides them."""
    pre, binned, img, _ = cases.oracle_run(sc)
    W, H = sc.width, sc.height
    rng = np.random.default_rng(seed)
    gx = (W + 15) // 16
    out = []
    tried = 0
    while len(out) < n_pixels and tried < 50 * n_pixels:
        tried += 1
        px, py = int(rng.integers(0, W)), int(rng.integers(0, H))
        n_c = int(img["n_contrib"][py, px])
        if n_c < 3:
            continue
        lo, hi = binned["ranges"][(py // 16) * gx + px // 16]
        ids = binned["point_list"][lo:hi][:n_c]
        contrib = []
        for g in ids:  # forward order; keep those the forward blended (forward.cu:418-426)
            xy, co = pre["means2D"][g], pre["conic_opacity"][g]
            dx, dy = F(xy[0]) - F(px), F(xy[1]) - F(py)
            power = F(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy
            if power > 0:
                continue
            alpha = min(F(0.99), co[3] * np.exp(power, dtype=F))
            if alpha < F(1.0 / 255.0):
                continue
            u = np.array([*pre["rgb"][g], pre["depths"][g], 1.0], dtype=F)  # colours, depth, alpha channel (u = 1)
            contrib.append((int(g), F(alpha), u))
        if len(contrib) >= 3:
            out.append(contrib[::-1])  # back to front, as the backward walks them
    return out


def _both(contrib, g_pix):
    ref, new = [], []
    acc = np.zeros(5, dtype=F); last_a = F(0); last_u = np.zeros(5, dtype=F)
    A = F(0)
    for _, a, u in contrib:
        acc = last_a * last_u + (F(1) - last_a) * acc          # reference, per channel
        ref.append(F(((u - acc) * g_pix).sum(dtype=F)))
        last_a, last_u = a, u
        s = F((u * g_pix).sum(dtype=F))                        # kernels, one scalar
        new.append(F(s - A))
        A = a * s + (F(1) - a) * A
    return np.array(ref, dtype=np.float64), np.array(new, dtype=np.float64)


@pytest.mark.parametrize("name", ["plumbing_s3", "street_small", "adversarial"])
def test_scalar_recurrence_matches_the_per_channel_recurrences(name):
    sc = cases.golden_cases()[name]
    rng = np.random.default_rng(5)
    per_gaussian_ref, per_gaussian_new = {}, {}
    worst = 0.0
    lists = _pixel_lists(sc, 160, seed=3)
    assert len(lists) >= 40
    for contrib in lists:
        g_pix = rng.standard_normal(5).astype(F) * np.array([1, 1, 1, 0.1, 1], dtype=F)
        ref, new = _both(contrib, g_pix)
        worst = max(worst, float(np.abs(ref - new).max() / max(np.abs(ref).max(), 1e-30)))
        for (g, _, _), r, n in zip(contrib, ref, new):
            per_gaussian_ref[g] = per_gaussian_ref.get(g, 0.0) + r
            per_gaussian_new[g] = per_gaussian_new.get(g, 0.0) + n
    a = np.array([per_gaussian_ref[g] for g in per_gaussian_ref])
    b = np.array([per_gaussian_new[g] for g in per_gaussian_ref])
    summed = float(np.abs(a - b).max() / np.abs(a).max())
    assert worst <= 2e-5, worst     # per contribution, relative to the pixel's largest term
    assert summed <= 2e-5, summed   # accumulated per Gaussian, relative to the tensor maximum
