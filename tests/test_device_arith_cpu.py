"""Integer shortcuts of the binning kernels, restated in numpy float32 / Python and checked over their whole domain
(the GPU parity tests only reach the rectangle sizes of real scenes):

* `emit_instances_kernel` (csrc/binning.cu): tile row = m / w through `trunc(float(m) * rcp_rn(float(w)))` and ONE
  correction step, used for m < 2^22 (an integer division otherwise);
* the warp-cooperative tile test of `preprocess_fwd_kernel` (csrc/preprocess_fwd.cu): `trunc((t + 0.5) * rcp_rn(w))`
  for t < 32, w <= 32 without a correction step;
* `nth_set_bit` (csrc/binning.cu): five popcount steps.
"""
import random

import numpy as np


def test_float_estimate_of_the_tile_row_is_exact_after_one_correction():
    rng = np.random.default_rng(0)
    for _ in range(20):
        w = rng.integers(1, 65536, size=200_000).astype(np.uint32)
        m = rng.integers(0, 1 << 22, size=200_000).astype(np.uint32)
        q = rng.integers(0, 1 << 22, size=200_000) // w  # multiples of w and their neighbours: the boundary cases
        mb = np.clip(q * w + rng.integers(-1, 2, size=200_000), 0, (1 << 22) - 1).astype(np.uint32)
        for mm in (m, mb):
            est = (mm.astype(np.float32) * (np.float32(1) / w.astype(np.float32))).astype(np.float32)
            ty = est.astype(np.uint32)  # cvt.rzi
            rem = mm.astype(np.int64) - ty.astype(np.int64) * w.astype(np.int64)
            ty = np.where(rem < 0, ty - 1, np.where(rem >= w, ty + 1, ty))
            assert np.array_equal(ty, mm // w)


def test_half_offset_estimate_needs_no_correction_for_small_rectangles():
    for w in range(1, 33):
        for t in range(32):
            est = np.float32((np.float32(t) + np.float32(0.5)) * (np.float32(1) / np.float32(w)))
            assert int(est) == t // w


def _nth_set_bit(mask, n):
    pos = 0
    c = bin(mask & 0xFFFF).count("1")
    if n >= c:
        pos, n = 16, n - c
    for width, m in ((8, 0xFF), (4, 0xF), (2, 0x3)):
        c = bin((mask >> pos) & m).count("1")
        if n >= c:
            pos, n = pos + width, n - c
    if n >= ((mask >> pos) & 1):
        pos += 1
    return pos


def test_nth_set_bit():
    random.seed(1)
    for _ in range(100_000):
        mask = random.getrandbits(32) or 1
        bits = [i for i in range(32) if mask >> i & 1]
        n = random.randrange(len(bits))
        assert _nth_set_bit(mask, n) == bits[n]
    for i in range(32):
        assert _nth_set_bit(1 << i, 0) == i and _nth_set_bit(0xFFFFFFFF, i) == i
